"""Times K3 (dct_stats_accumulate) and K0 (write calibration) on their own. GPU only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads, native
from digicamtoy_b200.stats import CubeStats

for name, kw in (('c4', workloads.c4_full_camera()), ('dark', workloads.c1_dark(0.))):
    g = NTraceGenerator(seed=1, **kw)
    E = 16384
    cube = g.generate(E)
    st = CubeStats(g)
    for _ in range(2):
        st.accumulate(cube)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        st.accumulate(cube)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('%s: K3 %.3f ms per %d events = %.0f GB/s read' % (name, ms, E, cube.numel() * 2 / ms / 1e6))
buf = torch.empty(1 << 30, dtype=torch.int16, device='cuda')   # 2 GiB
for _ in range(2):
    native.write_calibrate(buf.data_ptr(), buf.numel())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    native.write_calibrate(buf.data_ptr(), buf.numel())
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('K0 write-only: %.3f ms per 2 GiB = %.0f GB/s' % (ms, buf.numel() * 2 / ms / 1e6))
