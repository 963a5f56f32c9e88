#!/bin/bash
# Tuning build (SMX_TUNING=1) sweep of the pipelined kernel's schedule knobs at the headline configuration.
#   bash benchmarks/k1_sweep.sh > gpurun_out/k1_sweep.txt
# cost of an item in the static schedule = ca + (cb + cd * (factors - 1)) * k-steps + cc for a cold item  (SMX_FAST_COST=ca,cb,cc,cd)
export SMX_TUNING=1
python -m smolyax_b200._build > /dev/null 2>&1
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-standin --no-others --e2e-steps 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms  frac %.4f' % (d['ms_per_step'], d['roofline']['frac']))"; }
for cost in "3,1,2,0" "3,1,2,0.25" "3,1,2,0.5" "3,1,2,1" "2.5,1,2,0.25" "2.5,1,2,0.5" "3,0.75,2,0.5" "3,1,2.5,0.5" "3.5,1,2,0.5" "2,1,1,0.5" "3,1,1.5,0.5" "3,1,2,0"; do echo -n "cost $cost: "; SMX_FAST_COST=$cost run; done
