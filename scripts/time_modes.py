"""Times K1 in the constructor-reachable variants of SURVEY.md §8 f-3 on the C4 geometry:
sub_binning > 0 (on-grid NSB, scatter kernel) and voltage_drop=True, next to the default mode.  GPU only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads

E = 8192
base = workloads.c4_full_camera()
cases = [('continuous NSB (default)', {}),
         ('sub_binning=1', dict(sub_binning=1)),
         ('voltage_drop=True', dict(voltage_drop=True)),
         ('sub_binning=1, NSB 0.125 GHz', dict(sub_binning=1, nsb_rate=0.125)),
         ('continuous, NSB 0.125 GHz', dict(nsb_rate=0.125))]
for name, extra in cases:
    kw = dict(base); kw.update(extra)
    g = NTraceGenerator(seed=1, **kw)
    out = torch.empty((E, 1296, 50), dtype=torch.int16, device='cuda')
    for _ in range(2):
        g.generate(E, first_event=0, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3):
        g.generate(E, first_event=i * E, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('%-32s kernel %-7s %7.2f ms per %d events = %.3e samples/s = %6.1f GB/s  mean ADC %.2f' % (
        name, g.kernel, ms, E, out.numel() / ms * 1e3, out.numel() * 2 / ms / 1e6, out[:64].float().mean().item()))
