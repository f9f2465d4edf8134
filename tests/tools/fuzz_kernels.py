"""Randomised cross-checks of the generation kernels (GPU): for random configurations
   * sweep == scatter bit for bit (where the sweep kernel applies), grid == scatter in the on-grid mode,
   * tensor within 1 LSB of sweep on a small share of samples (where the tensor kernel applies),
   * fused statistics == separate statistics == NumPy,
   * chunked generation == one-shot generation.
Usage: fuzz_kernels.py [n_cases] [seed]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from digicamtoy_b200 import NTraceGenerator
from digicamtoy_b200.stats import CubeStats


def random_config(rs):
    dt = 4
    n_bins = int(rs.choice([2, 4, 10, 26, 50, 51, 64, 80, 92, 100, 130]))
    time_start = int(rs.choice([0, 0, 0, -20, 8])) // dt * dt
    P = int(rs.choice([1, 7, 33, 128, 129, 300]))
    rate_kind = rs.choice(['zero', 'low', 'mid', 'high', 'perpixel'])
    rate = {'zero': 0.0, 'low': 10 ** rs.uniform(-3, -1.5), 'mid': 10 ** rs.uniform(-1.5, -0.5),
            'high': 10 ** rs.uniform(-0.5, 0.5)}.get(rate_kind)
    if rate is None:
        rate = list(np.where(rs.rand(P) < 0.2, 0.0, 10 ** rs.uniform(-3, 0.3, size=P)))
    kw = dict(time_start=time_start, time_end=time_start + n_bins * dt, time_sampling=dt, n_pixels=P, nsb_rate=rate,
              crosstalk=float(rs.choice([0.0, 0.08, 0.3])), gain_nsb=bool(rs.rand() < 0.7), poisson=bool(rs.rand() < 0.8),
              sigma_e=float(rs.choice([0.0, 0.8, 1.25])), sigma_1=float(rs.choice([0.0, 0.8])), gain=float(rs.uniform(2, 8)),
              baseline=float(rs.choice([0, 10.4, 200, 500])), time_signal=float(time_start + rs.uniform(-10, n_bins * dt)),
              jitter=float(rs.choice([0.0, 0.3, 2.0])), n_photon=float(rs.choice([0, 0, 1, 10, 300, 3000])))
    if rs.rand() < 0.2:
        kw['sub_binning'] = 1                      # on-grid NSB mode: grid kernel against the scatter kernel
        if rs.rand() < 0.15 and not isinstance(rate, list):
            kw['nsb_rate'] = float(rs.uniform(3.0, 4.0))   # rate * dt >= 12: the generation-by-generation sampler
    return kw


def run_case(i, rs):
    kw = random_config(rs)
    seed = int(rs.randint(1 << 30))
    n_ev = int(rs.choice([1, 3, 17, 40]))
    tag = 'case %d: P=%d B=%d rate=%s npe=%g xt=%g' % (i, kw['n_pixels'], (kw['time_end'] - kw['time_start']) // 4,
                                                          'per-pixel' if isinstance(kw['nsb_rate'], list) else '%.4g' % kw['nsb_rate'],
                                                          kw['n_photon'], kw['crosstalk'])
    gens = {}
    for k in ('scatter', 'sweep', 'tensor', 'grid', 'auto'):
        try:
            gens[k] = NTraceGenerator(seed=seed, kernel=k, **kw)
        except ValueError as e:
            if 'no positive sample inside the simulated window' in str(e):
                return tag + '  skipped (window ends before the pulse template starts)'
            if k in ('scatter', 'auto'):
                raise AssertionError(tag + ': kernel %s refused the configuration: %s' % (k, e))
    cubes = {k: g.generate(n_ev, first_event=5) for k, g in gens.items()}
    torch.cuda.synchronize()
    ref = cubes['scatter']
    if 'sweep' in cubes:
        assert torch.equal(ref, cubes['sweep']), tag + ': sweep != scatter'
    if 'grid' in cubes:
        assert torch.equal(ref, cubes['grid']), tag + ': grid != scatter'
    if 'tensor' in cubes:
        d = (cubes['tensor'].int() - ref.int())
        g0 = gens['scatter']
        # expected rounding of the fp16 moments [LSB rms]: 1.1e-4 x gain x template peak x sqrt(overlapping pe)
        h_max = float(np.abs(g0.pulse_template(g0.params.knot_x)).max())
        err = 1.1e-4 * float(np.max(g0.gain)) * h_max * np.sqrt(max(1.0, float(np.max(g0.nsb_rate)) * 88.0))
        assert int(d.abs().max()) <= 1 + int(6 * err), tag + ': tensor off by %d (expected rms %.3g)' % (int(d.abs().max()), err)
        n_diff, n_all = int((d != 0).sum()), d.numel()
        assert n_diff <= 6 + 3.0 * n_all * min(1.0, 1.6 * err), tag + ': tensor differs on %d of %d (expected rms %.3g)' % (n_diff, n_all, err)
    auto = gens['auto']
    # chunked == one-shot (auto kernel)
    if n_ev >= 3:
        parts = torch.cat([auto.generate(1, first_event=5), auto.generate(n_ev - 1, first_event=6)])
        torch.cuda.synchronize()
        assert torch.equal(parts, cubes['auto']), tag + ': chunked != one-shot'
    # fused statistics == separate == numpy (histogram and sums)
    fused = CubeStats(auto, charge_lo=-500, n_charge_bins=3000)
    cube_f = fused.generate(n_ev, first_event=5)
    sep = CubeStats(auto, charge_lo=-500, n_charge_bins=3000)
    sep.accumulate(cubes['auto'])
    torch.cuda.synchronize()
    assert torch.equal(cube_f, cubes['auto']), tag + ': fused statistics changed the cube'
    a, b = fused.result(), sep.result()
    for key in ('adc_hist', 'bin_sum', 'bin_sumsq', 'pixel_sum', 'pixel_sumsq') + (('charge_hist',) if fused.with_charge else ()):
        assert np.array_equal(a[key], b[key]), tag + ': fused != separate in ' + key
    c = cubes['auto'].cpu().numpy().astype(np.int64)
    assert np.array_equal(a['adc_hist'], np.bincount(c.ravel(), minlength=4096)), tag + ': histogram != numpy'
    assert np.array_equal(a['bin_sum'], c.sum(axis=(0, 1))), tag + ': bin sums != numpy'
    for g in gens.values():
        g.close()
    return tag + '  kernels: ' + ','.join('%s' % k for k in gens) + ' auto=' + auto.kernel


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    t0 = time.time()
    for i in range(n):
        msg = run_case(i, rs)
        if i % 20 == 0:
            print(msg, flush=True)
    print('fuzz ok: %d cases in %.1f s' % (n, time.time() - t0))
