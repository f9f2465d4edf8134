"""Times the generation kernels across NSB rates (ac_dc.yml geometry) to place the auto-selection
thresholds (SWEEP_MIN_MEAN_PE in csrc/sweep.cuh, DCT_TC_MIN_MEAN_PE in csrc/tensor.cuh).  GPU only.
Usage: kernel_crossover.py [kernel ...]   (default: scatter sweep tensor)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads

E = 8192
KERNELS = tuple(sys.argv[1:]) or ('scatter', 'sweep', 'tensor')
for rate in (0.003, 0.01, 0.02, 0.04, 0.07, 0.125, 0.25, 0.4, 0.66, 1.0, 2.0):
    row = []
    for kernel in KERNELS:
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
    print('rate %.3f GHz  mean pe %.1f  ' % (rate, rate * 240) +
          '  '.join('%s %.2f ms' % (k, t) for k, t in zip(KERNELS, row)))
