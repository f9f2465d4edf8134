"""Times both generation kernels across NSB rates (ac_dc.yml geometry) to place the auto-selection
threshold (SWEEP_MIN_MEAN_PE in csrc/sweep.cuh).  GPU only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads

E = 8192
for rate in (0.003, 0.01, 0.02, 0.04, 0.07, 0.125, 0.25, 0.66, 1.0):
    row = []
    for kernel in ('scatter', 'sweep'):
        g = NTraceGenerator(seed=1, kernel=kernel, **workloads.c2_ac_dc(rate, 10.))
        out = torch.empty((E, 1296, 50), dtype=torch.int16, device='cuda')
        for _ in range(2):
            g.generate(E, first_event=0, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(3):
            g.generate(E, first_event=i * E, out=out)
        e1.record(); torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) / 3)
    print('rate %.3f GHz  mean pe %.1f  scatter %.2f ms  sweep %.2f ms  ratio %.2f' % (
        rate, rate * 240, row[0], row[1], row[0] / row[1]))
