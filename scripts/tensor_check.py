"""Tensor-core kernel (K1t) against the sweep kernel on the same seed: analog sums, ADC counts, timing."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from digicamtoy_b200 import NTraceGenerator, workloads

def main():
    n_small = int(os.environ.get('TC_EVENTS', 16))
    kw = workloads.c4_full_camera()
    out = {}
    gs = NTraceGenerator(seed=7, kernel='sweep', **kw)
    gt = NTraceGenerator(seed=7, kernel='tensor', **kw)
    adc_s, an_s = gs.generate_analog(n_small)
    torch.cuda.synchronize()
    adc_t, an_t = gt.generate_analog(n_small)
    torch.cuda.synchronize()
    d = (an_t.double() - an_s.double()) * 2.5       # LSB (gain_eff 2.5)
    print('analog diff [LSB]: mean %.3e rms %.3e max %.3e' % (d.mean().item(), d.pow(2).mean().sqrt().item(), d.abs().max().item()))
    per_bin = d.mean(dim=(0, 1)).cpu().numpy()
    print('per-bin mean diff [LSB]:', np.array2string(per_bin, precision=4, max_line_width=200))
    print('analog sweep mean %.4f tensor mean %.4f' % (an_s.mean().item(), an_t.mean().item()))
    dd = (adc_t.int() - adc_s.int())
    frac = (dd != 0).float().mean().item()
    print('adc: max |diff| %d, fraction differing %.3e, mean diff %.3e' % (dd.abs().max().item(), frac, dd.float().mean().item()))
    out.update(analog_rms_lsb=d.pow(2).mean().sqrt().item(), analog_mean_lsb=d.mean().item(),
               analog_max_lsb=d.abs().max().item(), adc_max_diff=int(dd.abs().max().item()), adc_frac_diff=frac)
    if os.environ.get('TC_TIME', '1') == '1':
        E = int(os.environ.get('TC_TIME_EVENTS', 16384))
        buf = torch.empty((E, 1296, 50), dtype=torch.int16, device='cuda')
        for name, g in (('sweep', gs), ('tensor', gt)):
            for _ in range(2):
                g.generate(E, first_event=0, out=buf)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(3):
                g.generate(E, first_event=(i + 1) * E, out=buf)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print('%s: %.2f ms per %d events  (%.1f GB/s)' % (name, ms, E, E * 1296 * 50 * 2 / ms / 1e6))
            out[name + '_ms'] = ms
    print(json.dumps(out))

if __name__ == '__main__':
    main()
