"""Tiny end-to-end exercise of every kernel for compute-sanitizer (memcheck).  GPU only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from digicamtoy_b200 import NTraceGenerator, native
from digicamtoy_b200.stats import CubeStats

def run(**kw):
    g = NTraceGenerator(seed=3, **kw)
    for kernel_events in (1, 7):
        c, t, q = g.generate(kernel_events, truth=True)
        st = CubeStats(g, charge_lo=-500, n_charge_bins=4000)
        st.accumulate(c)
        st.result()
    g.generate_host(5)
    g.trace_pe(2)
    g.close()

run(n_pixels=37, nsb_rate=0.5, n_photon=8., jitter=0.2)                                   # sweep, tile stats, ragged last CTA
run(n_pixels=130, nsb_rate=0.5, n_photon=8., kernel='scatter')                            # scatter poly
run(n_pixels=5, time_end=368, nsb_rate=0.003, n_photon=2, jitter=0.3)                    # 92 bins: warp stats
run(n_pixels=9, time_end=199, nsb_rate=0.3)                                               # odd bin count: generic paths
run(n_pixels=6, nsb_rate=0.3, sub_binning=1, n_photon=4)                                  # on-grid NSB
run(n_pixels=6, nsb_rate=0.5, pulse_shape_file='utils/pulse_SST-1M_AfterPreampLowGain.dat')   # 75 taps, 4 phases
run(n_pixels=4, time_start=0, time_end=100, time_sampling=3, nsb_rate=0.2)               # dt not a multiple of the knot step
buf = torch.empty(4099, dtype=torch.int16, device='cuda')
native.write_calibrate(buf.data_ptr(), buf.numel())
out = torch.empty(1000, dtype=torch.float32, device='cuda')
for k in native.SAMPLERS:
    native.test_sample(k, 0.3 if k != 'poisson' else 40.0, 5, 1, 1000, out.data_ptr())
torch.cuda.synchronize()
print('sanitize_small: all kernels ran')
