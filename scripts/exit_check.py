"""Does a process that used generator `kernel` exit cleanly?  Usage: exit_check.py kernel [close|noclose] [n_generators]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads
kernel = sys.argv[1]; mode = sys.argv[2] if len(sys.argv) > 2 else 'noclose'; n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
for i in range(n):
    kw = workloads.c2_ac_dc(1.0 if kernel == 'tensor' else 0.1, 10.)
    if kernel == 'grid':
        kw['sub_binning'] = 1
    g = NTraceGenerator(seed=1, kernel=kernel, **kw)
    out = g.generate(64)
    torch.cuda.synchronize()
    if mode == 'close':
        g.close()
print('done', kernel, mode, n, flush=True)
