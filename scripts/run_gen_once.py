"""One generation launch of a named kernel on the C4 (or dark) workload, for ncu.  Usage: run_gen_once.py kernel [c4|dark|acdc] [events]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from digicamtoy_b200 import NTraceGenerator, workloads

kernel = sys.argv[1] if len(sys.argv) > 1 else 'tensor'
wl = sys.argv[2] if len(sys.argv) > 2 else 'c4'
E = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
kw = {'c4': workloads.c4_full_camera, 'dark': lambda: workloads.c1_dark(n_photon=3),
      'acdc': lambda: workloads.c2_ac_dc(0.125, 10.)}[wl]()
g = NTraceGenerator(seed=7, kernel=kernel, **kw)
buf = torch.empty((E, kw['n_pixels'], g.params.n_bins), dtype=torch.int16, device='cuda')
for i in range(2):
    g.generate(E, first_event=i * E, out=buf)
torch.cuda.synchronize()
print('ok', g.kernel, float(buf.float().mean()))
