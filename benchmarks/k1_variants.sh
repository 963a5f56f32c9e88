#!/bin/bash
# A/B timing of the K1 variants at the headline config (cfg2, 1e6 points): one bench.py process per variant
# (the knobs are read once per process).  Usage: k1_variants.sh "ENV1=a ENV2=b" "ENV1=c" ...
# Writes one JSON line per variant to gpurun_out/k1_variants.jsonl and prints a summary.
out=gpurun_out/k1_variants.jsonl
mkdir -p gpurun_out
: > $out
for v in "$@"; do
  echo "== $v" >&2
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>>gpurun_out/k1_variants.err | sed "s/^{/{\"variant\": \"$v\", /" >> $out
done
python - <<'PY'
import json
for l in open("gpurun_out/k1_variants.jsonl"):
    d = json.loads(l)
    print(f'{d["variant"]:45s} {d["ms_per_step"]:.4f} ms  frac {d["roofline"]["frac"]:.4f}  parity {d["parity_max_rel_vs_oracle_first64"]:.2e}')
PY
