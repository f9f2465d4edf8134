"""Kernel time of a few fixed configurations (8192 events, C4 geometry), for comparing builds:
DCT_LIB_PATH=<other .so> python scripts/time_kernels.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads
E = 8192
base = workloads.c4_full_camera()
cases = [('c4 sweep', dict(kernel='sweep')), ('c4 grid (sub_binning=1)', dict(sub_binning=1)),
         ('c4 scatter', dict(kernel='scatter')), ('dark sweep', None)]
for name, extra in cases:
    kw = dict(base); 
    if extra is None:
        kw = workloads.c1_dark(0.)
    else:
        kw.update(extra)
    g = NTraceGenerator(seed=1, **kw)
    out = torch.empty((E, 1296, g.params.n_bins), dtype=torch.int16, device='cuda')
    for _ in range(2):
        g.generate(E, first_event=0, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(4):
        g.generate(E, first_event=i * E, out=out)
    e1.record(); torch.cuda.synchronize()
    print('%-26s kernel %-8s %7.2f ms per %d events' % (name, g.kernel, e0.elapsed_time(e1) / 4, E))
    g.close()
