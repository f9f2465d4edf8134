"""K1t against K1a per kept bin: rms of the analog difference and share of ADC samples that differ."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from digicamtoy_b200 import NTraceGenerator
ACDC = dict(time_start=0, crosstalk=0.08, gain_nsb=True, poisson=True, sigma_e=1.25, sigma_1=0.8, gain=5, baseline=280, time_signal=20, jitter=0.2)
for time_end, n in ((8, 4000), (200, 400)):
    kw = dict(n_pixels=70, nsb_rate=0.9, n_photon=0., time_end=time_end, time_sampling=4, **ACDC)
    gs, gt = NTraceGenerator(seed=11, kernel='sweep', **kw), NTraceGenerator(seed=11, kernel='tensor', **kw)
    a_s, an_s = gs.generate_analog(n); a_t, an_t = gt.generate_analog(n)
    torch.cuda.synchronize()
    d = (an_t.double() - an_s.double()) * float(gs.gain[0])
    rms = d.pow(2).mean(dim=(0, 1)).sqrt().cpu().numpy(); mean = d.mean(dim=(0, 1)).cpu().numpy()
    frac = (a_t != a_s).float().mean(dim=(0, 1)).cpu().numpy()
    print('time_end', time_end, 'samples per bin', n * 70)
    print('  rms  [LSB]', np.array2string(rms, precision=4, max_line_width=250))
    print('  mean [LSB]', np.array2string(mean, precision=5, max_line_width=250))
    print('  frac diff ', np.array2string(frac, precision=4, max_line_width=250))
