"""Static evidence from the built library (no GPU): per-kernel registers / stack / static shared memory from
`cuobjdump -res-usage`, and per-kernel counts of the SASS mnemonics that prove the TMA / mbarrier / FP64-tensor paths
(`cuobjdump -sass`).  Usage: python profiles/sass_summary.py [lib] > profiles/rNN_sass_summary.txt"""
import re, subprocess, sys, collections
lib = sys.argv[1] if len(sys.argv) > 1 else "smolyax_b200/libsmolyax_b200.so"
def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o.replace("void ", "").replace("(anonymous namespace)::", "")) for o in out]
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout.splitlines()
usage = {}
for i, l in enumerate(res):
    m = re.match(r"\s*Function (\S+):", l)
    if m:
        u = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", res[i + 1]))
        usage[m.group(1)] = u
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
KEYS = ["DMMA", "DFMA", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "MUFU.RCP64H", "LDL", "STL"]
counts = collections.defaultdict(collections.Counter)
cur = None
for l in sass:
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = m.group(1); continue
    if cur:
        for k in KEYS:
            if re.search(r"\b" + re.escape(k) + r"\b|\b" + re.escape(k) + r"[.\s]", l):
                counts[cur][k] += 1
names = sorted(usage)
dem = dict(zip(names, demangle(names)))
print(f"# {lib}: {len(names)} kernels (sm_100a)")
print(f"{'kernel':70s} {'REG':>4s} {'STACK':>5s} {'SHARED':>7s} " + " ".join(f"{k:>8s}" for k in KEYS))
for n in sorted(names, key=lambda n: dem[n]):
    u, c = usage[n], counts[n]
    print(f"{dem[n][:70]:70s} {u.get('REG','?'):>4s} {u.get('STACK','?'):>5s} {u.get('SHARED','?'):>7s} " + " ".join(f"{c[k]:8d}" for k in KEYS))
