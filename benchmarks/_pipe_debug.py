import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from smolyax_b200 import workloads
from smolyax_b200.interpolation import SmolyakBarycentricInterpolator
wl = workloads.CONFIGS['cfg2']
ip = SmolyakBarycentricInterpolator(node_gen=wl.generator(), k=wl.k(), t=wl.threshold(), d_out=1, batched_f=True)
layout = ip._assemble_compact(wl.target(), {})[0]
x = torch.rand((1_000_000, 1000), dtype=torch.float64, device='cuda') * 2 - 1
for env in (sys.argv[1:] or [""]):
    for kv in env.split():
        k, v = kv.split('='); os.environ[k] = v
    ip.set_layout(layout)
    for _ in range(3): ip(x)
    torch.cuda.synchronize()
    os.environ['SMX_PIPE_DEBUG'] = '1'
    print("==", env, file=sys.stderr, flush=True)
    ip(x); torch.cuda.synchronize()
    os.environ['SMX_PIPE_DEBUG'] = '0'
