"""K3 timing vs cube size (CUDA events), to cross-check ncu's per-launch duration."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads
from digicamtoy_b200.stats import CubeStats
g = NTraceGenerator(seed=1, **workloads.c4_full_camera())
for E in (1024, 4096, 16384):
    cube = g.generate(E)
    st = CubeStats(g)
    for _ in range(2):
        st.accumulate(cube)
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); st.accumulate(cube); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    print('E=%5d: K3 per launch %s ms  (min %.3f)' % (E, ' '.join('%.3f' % t for t in times), min(times)))
    # raw handle call without the python-side counter update
    f, i = st._f, st._i
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g._handle.stats_accumulate(cube.data_ptr(), E, adc_hist=st._ptr(i, st._slices_i['adc_hist']),
                               bin_sum=st._ptr(f, st._slices_f['bin_sum']), bin_sumsq=st._ptr(f, st._slices_f['bin_sumsq']),
                               pixel_sum=st._ptr(f, st._slices_f['pixel_sum']), pixel_sumsq=st._ptr(f, st._slices_f['pixel_sumsq']),
                               charge_hist=st._ptr(i, st._slices_i['charge_hist']), charge_lo=st.charge_lo,
                               n_charge_bins=st.n_charge_bins, moments=st._ptr(f, st._slices_f['moments']),
                               stream=torch.cuda.current_stream().cuda_stream)
    e1.record(); torch.cuda.synchronize()
    print('         raw call %.3f ms' % e0.elapsed_time(e1))
