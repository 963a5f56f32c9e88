import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from smolyax_b200 import workloads
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator
def timed(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); out=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); out.append(a.elapsed_time(b))
    return float(np.median(out))
for cfg, douts, n in (("cfg2", (1,2,3,4,8,16), 200_000), ("cfg4", (1,2,4,10), 100_000)):
    base = workloads.CONFIGS[cfg]
    for d_out in douts:
        wl = workloads.Workload(cfg, base.rule, base.d_in, d_out, base.n_target, n)
        ref = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=d_out, f=wl.target(), batched_f=True, dense=False)
        x = torch.from_numpy(wl.points(n, seed=1)).cuda()
        t_sparse = timed(lambda: ref(x))
        alt = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=d_out, dense=True)
        alt.set_layout(ref._layout)
        t_dense = timed(lambda: alt(x))
        err = float((ref(x) - alt(x)).abs().max())
        print(cfg, "d_out", d_out, "sparse ms %.3f dense ms %.3f  ratio %.2f  maxdiff %.1e" % (t_sparse, t_dense, t_sparse / t_dense, err), flush=True)
        del ref, alt
