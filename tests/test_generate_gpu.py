"""T2-T4: the stochastic kernels through the C ABI -- determinism / geometry independence,
scatter == sweep, per-sampler distributions, per-bin moments against the closed form, two-sample
tests against the CPU oracle, saturation, iterator semantics, on-device statistics."""
import numpy as np
import pytest
from scipy import stats as sps

from oracle import reference_port as orc

pytestmark = pytest.mark.gpu

ACDC = dict(time_start=0, time_end=200, time_sampling=4, crosstalk=0.08, gain_nsb=True, poisson=True,
            sigma_e=1.25, sigma_1=0.8, gain=5, baseline=280, time_signal=20, jitter=0.2)
DARK = dict(time_start=0, time_end=368, time_sampling=4, nsb_rate=0.003, crosstalk=0.08,
            gain_nsb=True, poisson=True, sigma_e=0.8, sigma_1=0.8, gain=5.8, baseline=200,
            time_signal=0., jitter=0.3)


def gen(**kw):
    from digicamtoy_b200 import NTraceGenerator
    return NTraceGenerator(**kw)


def cube(g, n, first_event=0, truth=False):
    import torch
    out = g.generate(n, first_event=first_event, truth=truth)
    torch.cuda.synchronize()
    if truth:
        return tuple(t.cpu().numpy() for t in out)
    return out.cpu().numpy()


# ------------------------------------------------------------------ determinism / geometry
@pytest.mark.parametrize('kernel', ['scatter', 'sweep'])
def test_chunking_and_offsets_do_not_change_bits(kernel):
    g = gen(n_pixels=37, nsb_rate=0.5, n_photon=8., seed=99, kernel=kernel, **ACDC)
    whole = cube(g, 40)
    parts = np.concatenate([cube(g, 7, 0), cube(g, 13, 7), cube(g, 20, 20)])
    assert np.array_equal(whole, parts)
    assert np.array_equal(cube(g, 5, first_event=17), whole[17:22])       # any range is regenerable
    g2 = gen(n_pixels=37, nsb_rate=0.5, n_photon=8., seed=99, kernel=kernel, **ACDC)
    assert np.array_equal(cube(g2, 40), whole)                             # same seed, new handle
    g3 = gen(n_pixels=37, nsb_rate=0.5, n_photon=8., seed=100, kernel=kernel, **ACDC)
    assert not np.array_equal(cube(g3, 40), whole)


@pytest.mark.parametrize('kernel', ['scatter', 'sweep'])
def test_event_index_beyond_32_bits(kernel):
    """The global event index is 56 bits wide in the Philox counter (low word + 24 high bits):
    a range that crosses 2^32 is the concatenation of its parts, and event 2^32 is not event 0."""
    g = gen(n_pixels=40, nsb_rate=0.3, n_photon=3., seed=3, kernel=kernel, **ACDC)
    edge = 1 << 32
    whole = cube(g, 6, first_event=edge - 3)
    assert np.array_equal(whole[:3], cube(g, 3, first_event=edge - 3))
    assert np.array_equal(whole[3:], cube(g, 3, first_event=edge))
    assert not np.array_equal(whole[3:], cube(g, 3, first_event=0))
    far = cube(g, 2, first_event=(1 << 55) + 5)
    assert far.std() > 0 and not np.array_equal(far, cube(g, 2, first_event=5))
    with pytest.raises(ValueError):
        g.generate(2, first_event=(1 << 56) - 1)


@pytest.mark.parametrize('rate,npe', [(1.0, 100.), (0.125, 10.), (0.04, 0.), (0.003, 1.)])
def test_sweep_and_scatter_kernels_are_bit_identical(rate, npe):
    a = gen(n_pixels=200, nsb_rate=rate, n_photon=npe, seed=5, kernel='scatter', **ACDC)
    b = gen(n_pixels=200, nsb_rate=rate, n_photon=npe, seed=5, kernel='sweep', **ACDC)
    assert a.kernel == 'scatter' and b.kernel == 'sweep'
    ca, ta, qa = cube(a, 24, truth=True)
    cb, tb, qb = cube(b, 24, truth=True)
    assert np.array_equal(ca, cb)
    assert np.array_equal(ta, tb) and np.array_equal(qa, qb)


def test_sweep_scatter_identical_with_per_pixel_rates_and_dark_window():
    kw = dict(DARK)
    kw['nsb_rate'] = np.concatenate([np.zeros(3), np.logspace(-3, 0.3, 61)])
    a = gen(n_pixels=64, n_photon=2, seed=1, kernel='scatter', **kw)
    b = gen(n_pixels=64, n_photon=2, seed=1, kernel='sweep', **kw)
    assert np.array_equal(cube(a, 30), cube(b, 30))


@pytest.mark.parametrize('kw', [
    dict(nsb_rate=1.0, n_photon=100., **ACDC),
    dict(nsb_rate=0.125, n_photon=0., **ACDC),
    dict(nsb_rate=3.5, n_photon=2., **ACDC),                               # Poisson mean 14 per bin: PTRS branch
    dict(n_photon=2, **{**DARK, 'sigma_1': 0., 'nsb_rate': np.concatenate([np.zeros(3), np.logspace(-3, 0.6, 61)])}),
    dict(nsb_rate=0.5, n_photon=5., **{**ACDC, 'crosstalk': 0.5}),           # long crosstalk chains: overflow blocks
])
def test_grid_and_scatter_kernels_are_bit_identical_for_on_grid_nsb(kw):
    """sub_binning > 0: the register-window kernel (default) against the generic one."""
    P = 64 if np.ndim(kw['nsb_rate']) else 200
    a = gen(n_pixels=P, sub_binning=1, seed=5, kernel='scatter', **kw)
    b = gen(n_pixels=P, sub_binning=1, seed=5, **kw)
    assert a.kernel == 'scatter' and b.kernel == 'grid'
    ca, ta, qa = cube(a, 24, truth=True)
    cb, tb, qb = cube(b, 24, truth=True)
    assert np.array_equal(ca, cb)
    assert np.array_equal(ta, tb) and np.array_equal(qa, qb)
    assert ca.std() > 0


def test_grid_kernel_falls_back_when_the_template_has_too_many_on_grid_taps():
    g = gen(n_pixels=8, sub_binning=1, nsb_rate=0.3, seed=1, time_start=0, time_end=100, time_sampling=2)
    assert g.kernel == 'scatter'                                           # 88 ns / 2 ns = 45 taps > 24
    assert cube(g, 4).shape == (4, 8, 50)
    g = gen(n_pixels=8, sub_binning=1, nsb_rate=0.3, seed=1,               # template starts at -8 ns
            pulse_shape_file='utils/pulse_SST-1M_AfterPreampLowGain.dat', **ACDC)
    assert g.kernel == 'scatter' and cube(g, 4).std() > 0


def test_host_path_equals_device_path():
    g = gen(n_pixels=1296, nsb_rate=0.125, n_photon=10., seed=3, **ACDC)
    dev, t, q = cube(g, 700, truth=True)           # > one 64 MiB chunk of the host pipeline
    adc, th, qh = g.generate_host(700, first_event=0)
    assert adc.dtype == np.uint16
    assert np.array_equal(adc.view(np.int16), dev)
    assert np.array_equal(th, t) and np.array_equal(qh, q)
    # pageable destination
    import ctypes
    buf = np.empty((700, 1296, 50), dtype=np.int16)
    g._handle.generate_host(g.seed, 0, 700, buf.ctypes.data)
    assert np.array_equal(buf, dev)


@pytest.mark.parametrize('transport', ['plain', 'packed12'])
@pytest.mark.parametrize('kernel,rate,n', [('sweep', 0.125, 1500), ('tensor', 1.0, 2300), ('tensor', 1.0, 200)])
def test_host_path_with_the_ramped_chunk_schedule(monkeypatch, transport, kernel, rate, n):
    """More than eight units of the host pipeline: chunks of 1, 2, 4 ... 4, 2, 1 units (for the persistent
    tensor kernel a unit is three rounds of its groups), device-side slot hand-over for a pinned
    destination, host-side for a pageable one -- the cube is the one of a single launch."""
    monkeypatch.setenv('DCT_HOST_TRANSPORT', transport)
    g = gen(n_pixels=1296, nsb_rate=rate, n_photon=10., seed=5, kernel=kernel, **ACDC)
    dev, t, q = cube(g, n, truth=True)
    adc, th, qh = g.generate_host(n, first_event=0)
    assert np.array_equal(adc.view(np.int16), dev)
    assert np.array_equal(th, t) and np.array_equal(qh, q)
    buf = np.empty((n, 1296, 50), dtype=np.int16)
    g._handle.generate_host(g.seed, 0, n, buf.ctypes.data)
    assert np.array_equal(buf, dev)
    g.close()


@pytest.mark.parametrize('kw,n', [
    (dict(n_pixels=7, time_start=0, time_end=204, kernel='scatter'), 333),      # odd P, odd B (51): byte-level tails
    (dict(n_pixels=129, time_start=0, time_end=204, kernel='scatter'), 41),
    (dict(n_pixels=33, time_start=0, time_end=200, kernel='sweep'), 1001),
    (dict(n_pixels=300, time_start=0, time_end=200, sub_binning=1), 97),        # grid kernel
    (dict(n_pixels=1, time_start=0, time_end=8, kernel='scatter'), 1),          # two samples in all
    (dict(n_pixels=1, time_start=0, time_end=4, kernel='scatter'), 3),          # one sample per event, odd total
])
def test_packed_host_transport_equals_the_device_cube(monkeypatch, kw, n):
    """dct_generate_host with DCT_HOST_TRANSPORT=packed12: the kernels write the chunk as a stream of
    12-bit samples (store_run_packed12), host threads widen it -- same cube as dct_generate, for every
    kernel, odd trace lengths and partial CTAs; pageable and pinned destinations; with saturating pulses."""
    monkeypatch.setenv('DCT_HOST_TRANSPORT', 'packed12')
    monkeypatch.setenv('DCT_HOST_THREADS', '3')
    base = dict(ACDC)
    base.update(kw)
    g = gen(nsb_rate=0.6, n_photon=2000., seed=8, **base)
    dev, t, q = cube(g, n, truth=True)
    if base['time_end'] >= 200:
        assert dev.max() == 4095                               # the rail is part of the test
    adc, th, qh = g.generate_host(n, first_event=0)
    assert np.array_equal(adc.view(np.int16), dev)
    assert np.array_equal(th, t) and np.array_equal(qh, q)
    buf = np.full(dev.shape, -1, dtype=np.int16)
    g._handle.generate_host(g.seed, 0, n, buf.ctypes.data)
    assert np.array_equal(buf, dev)
    g.close()


def test_packed_host_transport_is_not_used_when_the_adc_range_does_not_fit(monkeypatch):
    """saturate=None lets the ADC leave [0, 4095]: the 12-bit transport must stand aside."""
    monkeypatch.setenv('DCT_HOST_TRANSPORT', 'packed12')
    g = gen(n_pixels=50, nsb_rate=0.6, n_photon=3000., seed=8, saturate=None, **{**ACDC, 'baseline': -50.})
    dev = cube(g, 200)
    assert dev.max() > 4095 and dev.min() < 0
    adc = g.generate_host(200, first_event=0, truth=False)
    assert np.array_equal(adc.view(np.int16), dev)
    g.close()


# ------------------------------------------------------------------ the samplers, one by one
def _sample(kind, a=0.0, b=0.0, n=400000, seed=11):
    import torch
    from digicamtoy_b200 import native
    out = torch.empty(n, dtype=torch.float32, device='cuda')
    native.test_sample(kind, a, b, seed, n, out.data_ptr())
    torch.cuda.synchronize()
    return out.cpu().numpy().astype(np.float64)


def test_uniform_and_normal_samplers():
    u = _sample('uniform')
    assert u.min() >= 0 and u.max() < 1
    assert sps.kstest(u, 'uniform').pvalue > 1e-3
    z = _sample('normal')
    assert sps.kstest(z, 'norm').pvalue > 1e-3
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.var() - 1) < 0.01
    assert abs(sps.kurtosis(z)) < 0.05


def test_normal_sampler_reaches_beyond_six_sigma():
    """Box-Muller on a 23-bit uniform ends at sqrt(-2 ln 2^-24) = 5.77 sigma.  The radius uniform
    uses all 32 bits (2^-32 grid near 0): up to 6.77 sigma, and the right number of samples between."""
    import torch
    from digicamtoy_b200 import native
    r = _sample('normal_radius_top', n=1024)
    want = np.sqrt(-2 * np.log((np.arange(1024) + 0.5) * 2.0 ** -32))                  # 6.76 ... 5.52
    assert np.abs(r - want).max() < 2e-3 and r[0] > 6.75
    assert np.all(np.diff(r) <= 1e-6)
    n, chunks = 1 << 28, 8
    beyond55, beyond577, biggest = 0, 0, 0.0
    out = torch.empty(n, dtype=torch.float32, device='cuda')
    for c in range(chunks):
        native.test_sample('normal', 0, 0, 1000 + c, n, out.data_ptr())
        a = out.abs()
        beyond55 += int((a > 5.5).sum())
        beyond577 += int((a > 5.768).sum())
        biggest = max(biggest, float(a.max()))
    total = n * chunks                                               # 2.1e9 normals
    want55, want577 = total * 2 * sps.norm.sf(5.5), total * 2 * sps.norm.sf(5.768)     # 82 and 17
    assert abs(beyond55 - want55) < 5 * np.sqrt(want55), (beyond55, want55)
    assert abs(beyond577 - want577) < 5 * np.sqrt(want577) and beyond577 > 0, (beyond577, want577)
    assert 5.768 < biggest < 6.78


def test_exponential_gap_sampler():
    x = _sample('exp_gap', a=2.5)
    assert sps.kstest(x, 'expon', args=(0, 2.5)).pvalue > 1e-3


@pytest.mark.parametrize('xt', [0.0, 0.08, 0.3, 0.6])
def test_borel_sampler_matches_closed_form(xt):
    n = _sample('borel', a=xt)
    if xt == 0:
        assert np.all(n == 1)
        return
    pmf = orc.borel_pmf(xt, 200)
    kmax = int(np.searchsorted(-pmf * n.size, -8)) + 1       # bins with >= 8 expected
    obs = np.bincount(n.astype(int), minlength=kmax + 2)[1:kmax + 1].astype(float)
    exp = pmf[:kmax] * n.size
    obs = np.append(obs, n.size - obs.sum())
    exp = np.append(exp, n.size - exp.sum())
    chi2 = ((obs - exp) ** 2 / exp).sum()
    assert sps.chi2.sf(chi2, len(obs) - 1) > 1e-3, (chi2, len(obs))
    assert abs(n.mean() - 1 / (1 - xt)) < 5 * np.sqrt(xt / (1 - xt) ** 3 / n.size)


@pytest.mark.parametrize('lam', [0.012, 0.7, 4.0, 11.9, 12.1, 100.0, 3162.0])
def test_poisson_sampler(lam):
    k = _sample('poisson', a=lam, n=300000)
    assert abs(k.mean() - lam) < 5 * np.sqrt(lam / k.size)
    assert abs(k.var() / lam - 1) < 0.02
    lo, hi = int(max(0, lam - 6 * np.sqrt(lam) - 2)), int(lam + 6 * np.sqrt(lam) + 3)
    edges = np.arange(lo, hi + 1)
    exp = sps.poisson.pmf(edges, lam) * k.size
    keep = exp >= 8
    obs = np.array([(k == e).sum() for e in edges[keep]], dtype=float)
    obs = np.append(obs, k.size - obs.sum())
    exp = np.append(exp[keep], k.size - exp[keep].sum())
    ok = exp > 0
    chi2 = ((obs[ok] - exp[ok]) ** 2 / exp[ok]).sum()
    assert sps.chi2.sf(chi2, ok.sum() - 1) > 1e-4, (lam, chi2, ok.sum())


def test_inversion_samplers_stop_in_the_tail_for_the_largest_uniforms():
    """The float sum of a pmf can end a few 1e-8 below 1.  A search that only stops at
    ``u < cdf`` then runs to its cap for the largest uniforms (1 - 2^-23 ...): a 200-pe count or a
    256-pe crosstalk cluster once per ~1e7 draws.  Feed the samplers exactly those uniforms."""
    n = 64
    for kind in ('poisson_top', 'poisson_inv_top', 'poisson_table_top'):
        for lam in np.concatenate([np.linspace(0.004, 11.99, 300), [0.08, 0.32, 0.5, 4.0]]):
            v = _sample(kind, lam, 0, n)
            top = sps.poisson.isf(1e-10, lam) + 2
            assert v.max() <= top, (kind, lam, v.max(), top)
            assert v.min() >= sps.poisson.isf(2.0 ** -23 * n * 4, lam) - 1, (kind, lam, v.min())
            assert np.all(np.diff(v) <= 0)                         # smaller uniform, smaller count
    for lam in np.linspace(0.004, 11.99, 300):                     # the three forms agree draw by draw
        a, b, c = (_sample(k, lam, 0, n) for k in ('poisson_inv_top', 'poisson_table_top', 'poisson_top'))
        assert np.array_equal(a, b), lam
        assert np.abs(a - c).max() <= 3, lam                       # poisson() divides, the others multiply by 1/k:
                                                                   # at 1e-7 from the top the rounding shows
    for xt in np.concatenate([np.linspace(0.005, 0.5, 100), [0.08]]):
        v = _sample('borel_top', xt, 0, n)
        tail = 1 - np.cumsum(orc.borel_pmf(xt, 400))
        top = np.arange(1, 401)[np.argmax(tail < 1e-11)] + 2
        assert v.max() <= top, (xt, v.max(), top)
        assert np.all(np.diff(v) <= 0)


@pytest.mark.parametrize('n0,xt', [(1, 0.08), (7, 0.1), (100, 0.08), (3000, 0.3)])
def test_branching_total_is_generalised_poisson(n0, xt):
    s = _sample('branching', a=xt, b=n0, n=200000)
    mean, var = n0 / (1 - xt), n0 * xt / (1 - xt) ** 3
    assert abs(s.mean() - mean) < 5 * np.sqrt(var / s.size)
    assert abs(s.var() / var - 1) < 0.03
    if n0 == 1:   # a single tree is Borel
        pmf = orc.borel_pmf(xt, 50)
        obs = np.bincount(s.astype(int), minlength=8)[1:6]
        assert np.all(np.abs(obs - pmf[:5] * s.size) < 5 * np.sqrt(pmf[:5] * s.size) + 1)


@pytest.mark.parametrize('npe,xt', [(0.5, 0.0), (3.0, 0.08), (11.9, 0.08), (2.0, 0.4), (10.0, 0.3)])
def test_small_pulses_follow_the_generalised_poisson_law(npe, xt):
    """Pulses with n_photon < 12 take primaries + crosstalk progeny from one inversion in a
    generalised-Poisson table (reference: Poisson(n_photon) :236, one Galton-Watson tree per primary
    :319-324; closed form utils/pdf.py:37-42).  With sigma_1 = 0 the truth charge IS that total."""
    from scipy.special import gammaln
    g = gen(n_pixels=512, nsb_rate=0.0, n_photon=npe, crosstalk=xt, sigma_1=0.0, seed=33)
    _, _, q = cube(g, 800, truth=True)
    tot = q.ravel()
    assert np.array_equal(tot, np.round(tot)) and tot.min() >= 0
    k = np.arange(0, 200)
    with np.errstate(divide='ignore'):
        logp = np.log(npe) + (k - 1) * np.log(npe + k * xt) - (npe + k * xt) - gammaln(k + 1)
    pmf = np.exp(logp)
    assert abs(pmf.sum() - 1) < 1e-9
    obs = np.bincount(tot.astype(int), minlength=200)[:200].astype(float)
    keep = pmf * tot.size >= 10
    chi2 = ((obs[keep] - pmf[keep] * tot.size) ** 2 / (pmf[keep] * tot.size)).sum()
    rest_obs, rest_exp = obs[~keep].sum(), pmf[~keep].sum() * tot.size
    chi2 += (rest_obs - rest_exp) ** 2 / max(rest_exp, 1e-9) if rest_exp > 5 else 0.0
    assert sps.chi2.sf(chi2, keep.sum()) > 1e-3, (chi2, keep.sum())
    assert tot.mean() == pytest.approx(npe / (1 - xt), rel=4 * np.sqrt(npe / (1 - xt) ** 3 / tot.size) / (npe / (1 - xt)) + 1e-9)


# ------------------------------------------------------------------ moments vs closed form
def _moment_check(kw, n_events, kernel='auto', var_tol=0.03):
    P = kw['n_pixels']
    g = gen(seed=77, kernel=kernel, saturate=None, **kw)
    c = cube(g, n_events).astype(np.float64)                    # (E,P,B)
    cfg = orc.make_config(**kw)
    mean, var = orc.closed_form_moments(cfg)
    m_hat, v_hat = c.mean(axis=0), c.var(axis=0)
    z = (m_hat - mean) / np.sqrt(var / n_events)
    assert np.abs(z).max() < 6, 'per-(pixel,bin) mean off by %.1f sigma' % np.abs(z).max()
    assert abs(z.mean()) < 6 / np.sqrt(z.size)
    # pooled over pixels with identical parameters the variance is sharp
    if np.ptp(cfg.nsb_rate) == 0 and np.ptp(cfg.n_photon) == 0:
        ratio = v_hat.mean(axis=0) / var.mean(axis=0)
        assert np.abs(ratio - 1).max() < var_tol, ratio
    return g, c, cfg


def test_moments_dark_run():
    _moment_check(dict(n_pixels=64, n_photon=0, **DARK), 4000)


@pytest.mark.parametrize('rate', [0.04, 0.66, 1.0])
@pytest.mark.parametrize('kernel', ['scatter', 'sweep'])
def test_moments_nsb_levels(rate, kernel):
    _moment_check(dict(n_pixels=64, nsb_rate=rate, n_photon=0., **ACDC), 3000, kernel=kernel)


def test_moments_with_signal_pulse_and_edge_effect():
    g, c, cfg = _moment_check(dict(n_pixels=32, nsb_rate=1.0, n_photon=10., **ACDC), 6000, var_tol=0.05)
    m = c.mean(axis=(0, 1))
    # NSB only exists from -40 ns, the template is 88 ns long: early bins sit below true_baseline
    assert m[0] < g.true_baseline[0] - 1.0
    assert abs(m[-1] - g.true_baseline[0]) < 0.5


def test_moments_per_pixel_arrays_and_lowgain_template():
    kw = dict(n_pixels=5, nsb_rate=[0.0, 0.05, 0.2, 0.6, 1.2], crosstalk=[0.0, 0.05, 0.1, 0.2, 0.3],
              sigma_e=[0.5, 0.8, 1.0, 1.5, 2.0], sigma_1=[0.0, 0.4, 0.8, 1.0, 1.2],
              gain=[4., 5., 5.8, 6., 7.], baseline=[100, 200, 300, 400, 10],
              n_photon=[[0.], [3.], [30.], [300.], [1.]],
              time_signal=[[10.], [20.], [30.], [40.], [-20.]],
              jitter=[[0.], [0.5], [1.], [2.], [4.]], gain_nsb=0.7)
    _moment_check(kw, 20000)
    kw = dict(n_pixels=8, nsb_rate=0.5, gain_nsb=1 / 0.85, gain=5.8, baseline=500, sigma_e=0.8,
              sigma_1=0.8, crosstalk=0.08, time_end=200,
              pulse_shape_file='utils/pulse_SST-1M_AfterPreampLowGain.dat')
    _moment_check(kw, 8000)


def test_moments_sub_binning_and_dt2():
    _moment_check(dict(n_pixels=16, time_start=0, time_end=100, time_sampling=2, nsb_rate=0.2), 8000)
    # on-grid NSB: compare with the oracle's own sampling (no closed form for the grid process here)
    kw = dict(n_pixels=16, nsb_rate=0.3, sub_binning=1, n_photon=4)
    g = gen(seed=5, saturate=None, **kw)
    c = cube(g, 6000).astype(np.float64)
    cfg = orc.make_config(**kw)
    rng = np.random.RandomState(3)
    o = np.stack([orc.next_event(cfg, rng).adc_float for _ in range(400)])
    se = np.sqrt(o.var(axis=(0, 1)) / (o.shape[0] * o.shape[1]) + c.var(axis=(0, 1)) / (c.shape[0] * c.shape[1]))
    assert np.abs((c.mean(axis=(0, 1)) - o.mean(axis=(0, 1))) / se).max() < 5
    assert np.abs(c.var(axis=(0, 1)) / o.var(axis=(0, 1)) - 1).max() < 0.12


# ------------------------------------------------------------------ two-sample tests vs the oracle
def _two_sample_chi2_pvalue(a, b, min_expected=5.0):
    """Plain two-sample chi-square on integer samples: bins merged from both ends until every
    expected count of the smaller sample reaches `min_expected`; returns (p-value, dof)."""
    lo, hi = int(min(a.min(), b.min())), int(max(a.max(), b.max()))
    ha = np.bincount(a - lo, minlength=hi - lo + 1).astype(float)
    hb = np.bincount(b - lo, minlength=hi - lo + 1).astype(float)
    small = min(ha.sum(), hb.sum()) / (ha.sum() + hb.sum())
    groups, acc_a, acc_b = [], 0.0, 0.0
    for x, y in zip(ha, hb):
        acc_a += x; acc_b += y
        if (acc_a + acc_b) * small >= min_expected:
            groups.append((acc_a, acc_b)); acc_a = acc_b = 0.0
    if groups and (acc_a or acc_b):
        groups[-1] = (groups[-1][0] + acc_a, groups[-1][1] + acc_b)
    ga, gb = np.array(groups).T
    k1, k2 = np.sqrt(gb.sum() / ga.sum()), np.sqrt(ga.sum() / gb.sum())
    chi2 = ((k1 * ga - k2 * gb) ** 2 / (ga + gb)).sum()
    dof = len(groups) - 1
    return sps.chi2.sf(chi2, dof), dof


def one_sample_per_trace(adc, seed):
    """One randomly chosen bin of every trace: samples of different traces are independent, so a
    plain two-sample test applies (samples inside a trace share photo-electrons)."""
    flat = adc.reshape(-1, adc.shape[-1])
    pick = np.random.RandomState(seed).randint(0, adc.shape[-1], size=flat.shape[0])
    return flat[np.arange(flat.shape[0]), pick].astype(np.int64)


@pytest.mark.parametrize('kw,n_oracle', [
    (dict(n_pixels=48, nsb_rate=0.125, n_photon=10., **ACDC), 400),
    (dict(n_pixels=48, n_photon=2, **DARK), 400),
    (dict(n_pixels=48, nsb_rate=1.0, n_photon=100., **ACDC), 100),
])
def test_adc_and_charge_histograms_match_oracle(kw, n_oracle):
    g = gen(seed=2024, **kw)
    adc, t, q = cube(g, 4000, truth=True)
    cfg = orc.make_config(**kw)
    rng = np.random.RandomState(8)
    recs = [orc.next_event(cfg, rng) for _ in range(n_oracle)]
    o_adc = np.stack([orc.saturate(r.adc_float) for r in recs])
    o_q = np.stack([np.asarray(r.sig_amplitude, dtype=float) for r in recs])
    # ADC spectrum: one independent sample per trace, plain two-sample chi-square
    pval, dof = _two_sample_chi2_pvalue(one_sample_per_trace(adc, 1), one_sample_per_trace(o_adc, 2))
    assert dof >= 8 and pval > 1e-3, (pval, dof)
    # ... and the same for the sample at the pulse maximum region (bin of time_signal + 8..12 ns)
    b_peak = int((cfg.time_signal[0, 0] + 10.0 - kw['time_start']) // kw['time_sampling'])
    pval, dof = _two_sample_chi2_pvalue(adc[:, :, b_peak].ravel().astype(np.int64),
                                        o_adc[:, :, b_peak].ravel().astype(np.int64))
    assert pval > 1e-3, (pval, dof)
    # signal charge: KS two-sample (independent across traces)
    assert sps.ks_2samp(q.ravel(), o_q.ravel()).pvalue > 1e-3
    # signal time: uniform jitter around time_signal
    j = cfg.jitter[0, 0]
    assert t.min() >= cfg.time_signal[0, 0] - j / 2 - 1e-4 and t.max() <= cfg.time_signal[0, 0] + j / 2 + 1e-4
    assert sps.kstest((t.ravel() - cfg.time_signal[0, 0]) / j + 0.5, 'uniform').pvalue > 1e-3


# ------------------------------------------------------------------ saturation, edge cases
def test_twelve_bit_saturation_on_c4():
    P = 1296
    npe = [[float(10 ** (3.5 * i / (P - 1)))] for i in range(P)]
    g = gen(n_pixels=P, nsb_rate=1.0, n_photon=npe, seed=20261019, **ACDC)
    assert g.kernel == 'tensor'                                 # auto-selected at 240 pe / trace
    c = cube(g, 16)
    assert c.min() >= 0 and c.max() == 4095
    peak = c.max(axis=2).mean(axis=0)
    assert np.all(peak[-60:] == 4095) and np.all(peak[:600] < 4095)
    # unsaturated build shows what was clipped
    g2 = gen(n_pixels=P, nsb_rate=1.0, n_photon=npe, seed=20261019, saturate=None, **ACDC)
    c2 = cube(g2, 16)
    assert c2.max() > 4095 and np.array_equal(np.clip(c2, 0, 4095), c)


def test_negative_values_clamp_to_zero():
    g = gen(n_pixels=8, baseline=0, nsb_rate=0.01, sigma_e=2.0, n_photon=0, seed=4)
    c = cube(g, 200)
    assert c.min() == 0 and (c == 0).mean() > 0.2


def test_degenerate_configs():
    # nothing random at all: every sample equals round(baseline)
    g = gen(n_pixels=3, nsb_rate=0, crosstalk=0, sigma_e=0, sigma_1=0, n_photon=0, baseline=[10, 20.4, 30.6])
    c = cube(g, 5)
    assert np.array_equal(c[:, 0], np.full((5, 50), 10)) and np.all(c[:, 1] == 20) and np.all(c[:, 2] == 31)
    # deterministic pulse: poisson off, no crosstalk, no smearing -> analytic trace, incl. t on a sample
    g = gen(n_pixels=2, nsb_rate=0, crosstalk=0, sigma_e=0, sigma_1=0, n_photon=500, poisson=False,
            gain_nsb=False, gain=5.0, baseline=100, time_signal=20, jitter=0)
    c = cube(g, 3)
    t = np.arange(0, 200, 4)
    want = np.round(100 + 5.0 * 500 * g.get_pulse_shape(t, t_0=20)).astype(int)
    assert np.abs(c[0, 0] - want).max() <= 1 and (c[0, 0] != want).sum() <= 1
    assert want[5] > 100                                          # h(0) > 0 counted at t == time_signal
    # zero events / one pixel / odd bin count
    assert cube(g, 0).shape == (0, 2, 50)
    g = gen(n_pixels=1, time_end=199, nsb_rate=0.3, seed=1)
    assert cube(g, 3).shape == (3, 1, 50)
    g = gen(n_pixels=5, time_start=0, time_end=102, time_sampling=6, nsb_rate=0.3, seed=1)   # dt not a multiple of the knot step? 6/0.5=12 phases
    assert cube(g, 3).shape[2] == g.params.n_bins


# ------------------------------------------------------------------ iterator semantics (reference tests)
def test_n_events_like_reference_test():
    for n in (0, 1, 10, 100):                                     # test_NTraceGenerator.py:4-13
        toy = gen(nsb_rate=0.000001, n_events=n, n_pixels=16)
        assert len([e for e in toy]) == n


def test_sampling_like_reference_test():
    param = {'time_start': 0, 'time_end': 100, 'time_sampling': 2}   # test.py:7-13
    toy = gen(n_pixels=4, **param)
    toy.next()
    assert toy.adc_count.shape[-1] == (param['time_end'] - param['time_start']) // param['time_sampling']


def test_iterator_surface():
    toy = gen(n_pixels=3, baseline=[0, 500, 0], n_photon=[[100, 133], [60, 60, 60], [100, 133]],
              time_signal=[[0, 100], [20, 50, 100], [0, 100]], gain=[10, 10, 29], poisson=True,
              n_events=5, seed=1, batch_events=2, pulse_shape_file='utils/pulse_SST-1M_pixel_0.dat',
              some_unknown_option=1)                              # example.py:8-12; extra kwargs ignored
    assert iter(toy) is toy and toy.count == -1
    assert toy.adc_count.shape == (3, 60) and not toy.adc_count.any()    # before the first event (:199)
    seen = []
    for ev in toy:
        assert ev is toy
        assert ev.adc_count.shape == (3, 50) and ev.cherenkov_time.shape == (3, 3)
        assert ev.cherenkov_photon.shape == (3, 3) and ev.cherenkov_photon[0, 2] == 0
        assert ev.sampling_bins.shape == (50,) and ev.sampling_bins[0] == 0
        seen.append((ev.count, np.array(ev.adc_count)))
    assert [s[0] for s in seen] == [0, 1, 2, 3, 4]
    bulk = cube(gen(n_pixels=3, baseline=[0, 500, 0], n_photon=[[100, 133], [60, 60, 60], [100, 133]],
                    time_signal=[[0, 100], [20, 50, 100], [0, 100]], gain=[10, 10, 29], seed=1), 5)
    assert all(np.array_equal(s[1], bulk[i]) for i, s in enumerate(seen))   # batching is invisible
    with pytest.raises(AttributeError):
        toy.no_such_attribute
    assert toy.true_baseline.shape == (3,) and toy.tau > 17


# ------------------------------------------------------------------ K3 statistics kernel
def test_stats_kernel_matches_numpy():
    import torch
    g = gen(n_pixels=50, nsb_rate=0.3, n_photon=5., seed=9, **ACDC)
    out = g.generate(300)
    from digicamtoy_b200.stats import CubeStats
    st = CubeStats(g, charge_lo=-500, n_charge_bins=4000)
    st.accumulate(out[:100])
    st.accumulate(out[100:])
    r = st.result()
    c = out.cpu().numpy().astype(np.int64)
    assert np.array_equal(r['adc_hist'], np.bincount(c.ravel(), minlength=4096))
    assert np.allclose(r['bin_sum'], c.sum(axis=(0, 1))) and np.allclose(r['bin_sumsq'], (c ** 2).sum(axis=(0, 1)))
    assert np.allclose(r['pixel_sum'], c.sum(axis=(0, 2))) and np.allclose(r['pixel_sumsq'], (c ** 2).sum(axis=(0, 2)))
    charge = c.sum(axis=2) - 50 * 280
    assert np.array_equal(r['charge_hist'], np.bincount(np.clip(charge.ravel() + 500, 0, 3999), minlength=4000))
    x = c.ravel().astype(np.float64)
    assert np.allclose(r['moments'], [x.sum(), (x ** 2).sum(), (x ** 3).sum(), (x ** 4).sum()], rtol=1e-12)
    mean, mean_err, std, std_err = st.moments_nsb()
    assert mean == pytest.approx(x.mean()) and std == pytest.approx(x.std())


def _check_stats_against_numpy(g, out, charge_lo, n_charge_bins, pieces=(None,)):
    from digicamtoy_b200.stats import CubeStats
    st = CubeStats(g, charge_lo=charge_lo, n_charge_bins=n_charge_bins)
    edges = [0] + [p for p in pieces if p] + [out.shape[0]]
    for lo, hi in zip(edges[:-1], edges[1:]):
        st.accumulate(out[lo:hi])
    r = st.result()
    c = out.cpu().numpy().astype(np.int64)
    P, B = c.shape[1], c.shape[2]
    assert np.array_equal(r['adc_hist'], np.bincount(c.ravel(), minlength=4096))
    assert np.allclose(r['bin_sum'], c.sum(axis=(0, 1)), rtol=1e-13)
    assert np.allclose(r['bin_sumsq'], (c ** 2).sum(axis=(0, 1)), rtol=1e-13)
    assert np.allclose(r['pixel_sum'], c.sum(axis=(0, 2)), rtol=1e-13)
    assert np.allclose(r['pixel_sumsq'], (c ** 2).sum(axis=(0, 2)), rtol=1e-13)
    base = np.rint(np.broadcast_to(np.asarray(g.baseline, dtype=np.float32), (P,))).astype(np.int64)
    charge = c.sum(axis=2) - B * base[None, :]
    want = np.bincount(np.clip(charge.ravel() - charge_lo, 0, n_charge_bins - 1), minlength=n_charge_bins)
    if B <= 128:
        assert np.array_equal(r['charge_hist'], want)
    x = c.ravel().astype(np.float64)
    assert np.allclose(r['moments'], [x.sum(), (x ** 2).sum(), (x ** 3).sum(), (x ** 4).sum()], rtol=1e-12)
    # the host-side twin used by the gloo test (CubeStats.accumulate_numpy) gives the same buffers
    if g.saturate == (0, 4095):
        twin = CubeStats.from_shapes(P, B, saturate=g.saturate, baseline=g.baseline, charge_lo=charge_lo,
                                     n_charge_bins=n_charge_bins)
        twin.accumulate_numpy(c)
        t = twin.result()
        for key in ('adc_hist', 'bin_sum', 'bin_sumsq', 'pixel_sum', 'pixel_sumsq') + (('charge_hist',) if B <= 128 else ()):
            assert np.array_equal(t[key], r[key]), key


@pytest.mark.parametrize('name,kw,n_events', [
    ('tile-full-range', dict(n_pixels=200, nsb_rate=0.3, n_photon=5., **ACDC), 150),
    ('tile-compact-dim', dict(n_pixels=130, n_photon=0., **DARK), 200),                          # B = 92
    ('tile-compact-bright', dict(n_pixels=130, n_photon=[[3000.]] * 65 + [[40.]] * 65,            # saturating half
                                 **{**DARK, 'time_signal': 100.}), 200),
    ('tile-compact-128', dict(n_pixels=33, nsb_rate=0.2, n_photon=300., **{**ACDC, 'time_end': 512}), 120),
    ('tile-70-bins', dict(n_pixels=150, nsb_rate=0.2, n_photon=30., **{**ACDC, 'time_end': 280}), 90),    # 35 words: half-warps share 3 columns
    ('tile-96-bins', dict(n_pixels=77, nsb_rate=0.2, n_photon=30., **{**ACDC, 'time_end': 384}), 90),     # 48 words: ... 16 columns
    ('tile-98-bins', dict(n_pixels=77, nsb_rate=0.2, n_photon=30., **{**ACDC, 'time_end': 392}), 90),     # 49 words: one lane per column again
    ('generic-long', dict(n_pixels=20, nsb_rate=0.2, n_photon=30., **{**ACDC, 'time_end': 800}), 60),
    ('generic-odd', dict(n_pixels=20, nsb_rate=0.2, n_photon=30., **{**ACDC, 'time_end': 204}), 60),
])
def test_stats_kernel_every_dispatch_path(name, kw, n_events):
    """K3 has three kernels (tile: full-range / compact layout, generic); all must equal NumPy,
    including samples outside the histogram windows, saturated ones and clamped charges."""
    g = gen(seed=21, **kw)
    out = g.generate(n_events)
    _check_stats_against_numpy(g, out, charge_lo=-300, n_charge_bins=2500, pieces=(n_events // 3,))


def test_stats_kernel_synthetic_extremes():
    """Cubes that no generator config produces: every value of the 12-bit range, in long traces
    (compact layout: most samples lie outside the CTA histogram window)."""
    import torch
    g = gen(n_pixels=130, n_photon=0., seed=1, **DARK)
    B = g.generate(1).shape[2]
    assert B == 92
    rs = np.random.RandomState(3)
    c = rs.randint(0, 4096, size=(64, 130, B)).astype(np.int16)
    c[5] = 0
    c[6] = 4095
    c[7, :, ::2] = 4095
    out = torch.from_numpy(c).cuda()
    _check_stats_against_numpy(g, out, charge_lo=-1000, n_charge_bins=5000)


@pytest.mark.parametrize('geometry', ['acdc-50-bins', 'dark-92-bins'])
def test_stats_sums_of_squares_do_not_wrap_on_a_fully_saturated_wide_range_cube(geometry):
    """the tile kernel keeps a trace's sum of squares in 32 bits.  With saturate=(0, 8191) a
    fully saturated 92-bin trace holds 92 x 8191^2 = 6.2e9 > 2^32: the host gate must send that cube to
    the generic kernel; the 50-bin one (3.4e9) still fits the fast path.  Either way: NumPy's numbers."""
    import torch
    kw = dict(n_pixels=130, nsb_rate=0.01, seed=1, saturate=(0, 8191))
    kw.update(ACDC if geometry.startswith('acdc') else DARK)
    if 'nsb_rate' in DARK and not geometry.startswith('acdc'):
        kw['nsb_rate'] = 0.003
    g = gen(**kw)
    B = g.params.n_bins
    c = np.full((40, 130, B), 8191, dtype=np.int16)
    c[::3, ::2] = 8000
    out = torch.from_numpy(c).cuda()
    from digicamtoy_b200.stats import CubeStats
    st = CubeStats(g, charge_lo=0, n_charge_bins=1000)
    st.accumulate(out)
    r = st.result()
    c64 = c.astype(np.int64)
    assert np.array_equal(r['pixel_sumsq'], (c64 ** 2).sum(axis=(0, 2)))
    assert np.array_equal(r['bin_sumsq'], (c64 ** 2).sum(axis=(0, 1)))
    assert np.array_equal(r['pixel_sum'], c64.sum(axis=(0, 2)))
    assert np.array_equal(r['adc_hist'], np.bincount(c64.ravel(), minlength=8192))
    assert r['charge_hist'][-1] == 40 * 130                       # every charge beyond the last bin


FUSED_CASES = [
    ('sweep-acdc', dict(n_pixels=200, nsb_rate=0.3, n_photon=5., **ACDC), 150, 'sweep'),
    ('sweep-dark-92', dict(n_pixels=130, n_photon=2., **DARK), 200, 'sweep'),
    ('sweep-bright', dict(n_pixels=130, n_photon=[[3000.]] * 65 + [[40.]] * 65, **{**DARK, 'time_signal': 100.}), 100, 'sweep'),
    ('scatter-odd-bins', dict(n_pixels=20, nsb_rate=0.2, n_photon=30., **{**ACDC, 'time_end': 204}), 60, 'scatter'),
    ('scatter-lowgain', dict(n_pixels=77, nsb_rate=0.5, gain_nsb=1 / 0.85, baseline=500,
                             pulse_shape_file='utils/pulse_SST-1M_AfterPreampLowGain.dat'), 80, 'scatter'),
    ('grid', dict(n_pixels=100, nsb_rate=0.5, n_photon=5., sub_binning=1, **ACDC), 100, 'grid'),
    ('tensor', dict(n_pixels=150, nsb_rate=1.0, n_photon=50., **ACDC), 90, 'tensor'),
    ('negative-rail', dict(n_pixels=40, baseline=1, nsb_rate=0.01, sigma_e=2.0, n_photon=0), 120, 'sweep'),
]


@pytest.mark.parametrize('name,kw,n_events,kernel', FUSED_CASES, ids=[c[0] for c in FUSED_CASES])
def test_fused_statistics_equal_the_separate_kernel_and_numpy(name, kw, n_events, kernel):
    """dct_generate_stats (statistics taken inside the generation kernel) against
    dct_generate + dct_stats_accumulate on the same events, and both against NumPy: same cube, same
    histogram / sums bit for bit (integers < 2^53 in float64), power sums to rounding."""
    import torch
    from digicamtoy_b200.stats import CubeStats
    g = gen(seed=21, kernel=kernel, **kw)
    assert g.kernel == kernel
    fused = CubeStats(g, charge_lo=-300, n_charge_bins=2500)
    first = 0
    cubes = []
    for n in (n_events // 3, n_events - n_events // 3):        # buffers accumulate over calls
        cubes.append(fused.generate(n, first_event=first))
        first += n
    cube_f = torch.cat(cubes)
    plain = g.generate(n_events, first_event=0)
    torch.cuda.synchronize()
    assert torch.equal(cube_f, plain)                           # statistics do not touch the cube
    sep = CubeStats(g, charge_lo=-300, n_charge_bins=2500)
    sep.accumulate(plain)
    a, b = fused.result(), sep.result()
    assert a['n_events'] == b['n_events'] == n_events
    for key in ('adc_hist', 'charge_hist', 'bin_sum', 'bin_sumsq', 'pixel_sum', 'pixel_sumsq'):
        assert np.array_equal(a[key], b[key]), key
    assert np.allclose(a['moments'], b['moments'], rtol=1e-12)
    _check_stats_against_numpy(g, plain, charge_lo=-300, n_charge_bins=2500)


def test_fused_statistics_need_their_buffers():
    import torch
    g = gen(n_pixels=8, nsb_rate=0.1, seed=1)
    out = torch.empty((2, 8, 50), dtype=torch.int16, device='cuda')
    with pytest.raises(ValueError):
        g._handle.generate_stats(g.seed, 0, 2, out.data_ptr())


def test_write_calibrate_kernel():
    import torch
    from digicamtoy_b200 import native
    buf = torch.full(((1 << 20) + 3,), -1, dtype=torch.int16, device='cuda')
    native.write_calibrate(buf.data_ptr(), buf.numel())
    torch.cuda.synchronize()
    assert int((buf < 0).sum()) == 0


# ------------------------------------------------------------------ closing the loop: GPU pe lists -> reference arithmetic
@pytest.mark.parametrize('kernel', ['scatter', 'sweep'])
@pytest.mark.parametrize('rate,npe', [(1.0, 100.), (0.125, 10.), (0.003, 2.)])
def test_reference_arithmetic_on_gpu_drawn_photoelectrons_reproduces_gpu_adc(kernel, rate, npe):
    """The stochastic kernels never store their photo-electrons; trace_pe re-derives them from the
    same Philox counters.  Feeding those lists, the pulse truth and the noise through the ORACLE's
    shaping + digitisation (float64, the reference's own arithmetic, reference :350-381) must give
    the ADC counts the GPU wrote: every stage of K1 checked against the reference on identical
    draws.  float32 accumulation on the GPU -> at most 1 LSB on rounding ties."""
    P = 160
    kw = dict(n_pixels=P, nsb_rate=rate, n_photon=npe, **ACDC)
    g = gen(seed=321, kernel=kernel, saturate=None, **kw)
    n_ev = 6
    adc, t, q = cube(g, n_ev, truth=True)
    pe = g.trace_pe(n_ev, first_event=0)
    cfg = orc.make_config(**kw)
    assert pe['count'].mean() == pytest.approx(rate * 240, rel=0.05)
    n_diff = 0
    for e in range(n_ev):
        want = orc.shape_and_digitise(cfg, pe['time'][e], pe['amplitude'][e].astype(np.float64),
                                      t[e].astype(np.float64), q[e].astype(np.float64),
                                      pe['noise'][e].astype(np.float64))
        d = np.abs(adc[e].astype(np.int64) - want.astype(np.int64))
        assert d.max() <= 1, d.max()
        n_diff += int((d != 0).sum())
    assert n_diff <= 2e-3 * adc.size, n_diff


def closed_loop_against_oracle(g, kw, n_ev, max_frac, first_event=0):
    """GPU ADC of `n_ev` events against the oracle's float64 shaping + digitisation of the pe lists,
    pulse truth and noise the same counters give (trace_pe).  With saturation on, in-range samples
    must agree to 1 LSB, out-of-range ones must sit on the rail, and the number of clamped samples
    must equal the number of oracle values outside the range.  Returns (n_diff, n_clamped)."""
    adc, t, q = cube(g, n_ev, first_event=first_event, truth=True)
    pe = g.trace_pe(n_ev, first_event=first_event)
    cfg = orc.make_config(**kw)
    lo, hi = g.saturate
    n_diff = n_clamped = n_rail = 0
    for e in range(n_ev):
        want = orc.shape_and_digitise(cfg, pe['time'][e], pe['amplitude'][e].astype(np.float64),
                                      t[e].astype(np.float64), q[e].astype(np.float64),
                                      pe['noise'][e].astype(np.float64)).astype(np.int64)
        got = adc[e].astype(np.int64)
        inside = (want >= lo) & (want <= hi)
        d = np.abs(got - want)[inside]
        assert d.max() <= 1, d.max()
        n_diff += int((d != 0).sum())
        out = ~inside
        n_clamped += int(out.sum())
        # a value the oracle puts 1 LSB outside the range may be 1 LSB inside here (float32 tie)
        assert np.all(np.abs(got[out] - np.clip(want[out], lo, hi)) <= 1)
        n_rail += int(((got == lo) | (got == hi)).sum())
    assert n_diff <= max_frac * adc.size, (n_diff, adc.size)
    return n_diff, n_clamped, n_rail


@pytest.mark.parametrize('kernel', ['sweep', 'tensor'])
def test_closed_loop_on_the_full_c4_camera_including_saturated_pixels(kernel):
    """the closed loop on C4's own parameter set -- all 1296 pixels, 1 GHz NSB, pulses
    of 1..3162 pe, 12-bit saturation (the brightest ~10 % of the pixels clip)."""
    from digicamtoy_b200 import workloads
    kw = workloads.c4_full_camera()
    g = gen(seed=workloads.C4_SEED, kernel=kernel, **kw)
    assert g.kernel == kernel and g.saturate == (0, 4095)
    # the tensor kernel adds ~2.6e-3 LSB rms of fp16 moment rounding to the analog sums: the share of
    # samples that round the other way grows from ~1e-4 to ~2e-3
    n_diff, n_clamped, n_rail = closed_loop_against_oracle(g, kw, 2, 2e-3 if kernel == 'sweep' else 5e-3, first_event=5)
    assert n_clamped > 500                                    # saturation really is exercised
    assert abs(n_rail - n_clamped) <= 2 + 0.002 * n_clamped   # clamp count == oracle's out-of-range count (ties aside)


def test_closed_loop_on_dark_yaml():
    """... and on dark.yml's 92 bins (3 MHz dark counts, 3 pe pulses at t = 0, jitter 0.3 ns)."""
    from digicamtoy_b200 import workloads
    kw = workloads.c1_dark(n_photon=3)
    g = gen(seed=12, **kw)
    n_diff, n_clamped, _ = closed_loop_against_oracle(g, kw, 3, 2e-3)
    assert n_clamped == 0


@pytest.mark.parametrize('kernel', ['scatter', 'grid'])
@pytest.mark.parametrize('rate,xt', [(1.0, 0.08), (0.125, 0.08), (0.4, 0.5), (3.5, 0.08)])
def test_reference_arithmetic_reproduces_gpu_adc_in_on_grid_mode(kernel, rate, xt):
    """Same closed loop for sub_binning > 0 (reference :259-267): the pe sit on the simulated samples,
    dt apart from time_start - 40 ns on; counts per bin are Poisson(rate dt), clusters carry their
    crosstalk progeny (long chains and Poisson means >= 12 take the overflow draws)."""
    P = 96
    kw = dict(n_pixels=P, nsb_rate=rate, n_photon=5., sub_binning=1, **{**ACDC, 'crosstalk': xt})
    g = gen(seed=77, kernel=kernel, saturate=None, **kw)
    assert g.kernel == kernel
    n_ev = 6
    adc, t, q = cube(g, n_ev, truth=True)
    pe = g.trace_pe(n_ev, first_event=0)
    cfg = orc.make_config(**kw)
    # occupied bins: P(n > 0) = 1 - exp(-rate dt); every pe time is on the grid
    occupied = pe['count'].mean() / 60.0
    assert occupied == pytest.approx(1 - np.exp(-rate * 4), abs=0.02)
    n_diff = 0
    amp_sum = 0.0
    for e in range(n_ev):
        tt = pe['time'][e]
        for pix in range(P):
            k = pe['count'][e, pix]
            on_grid = (tt[pix, :k] + 40.0) / 4.0
            assert np.array_equal(on_grid, np.round(on_grid))
            amp_sum += pe['amplitude'][e, pix, :k].sum()
        want = orc.shape_and_digitise(cfg, tt, pe['amplitude'][e].astype(np.float64),
                                      t[e].astype(np.float64), q[e].astype(np.float64),
                                      pe['noise'][e].astype(np.float64))
        d = np.abs(adc[e].astype(np.int64) - want.astype(np.int64))
        assert d.max() <= 1, d.max()
        n_diff += int((d != 0).sum())
    assert n_diff <= 2e-3 * adc.size, n_diff
    # total charge per trace: rate dt pe per bin, each 1 / (1 - xt) after crosstalk
    want_sum = n_ev * P * 60 * rate * 4 / (1 - xt)
    assert amp_sum == pytest.approx(want_sum, rel=0.04 if xt < 0.3 else 0.08)


@pytest.mark.parametrize('rate,xt', [(1.0, 0.08), (0.125, 0.08), (0.5, 0.4)])
def test_on_grid_mode_high_statistics_against_closed_form(rate, xt):
    """sub_binning > 0 at 1.3e7 traces: per-bin mean and variance of the compound Poisson sum
    V_b = g sum_m h(m dt) S_(b-m), S = clusters of Poisson(rate dt) primaries (:259-267, :306-312)."""
    from digicamtoy_b200.stats import CubeStats
    P, E, chunk = 64, 200_000, 20_000
    kw = dict(n_pixels=P, nsb_rate=rate, n_photon=0., sub_binning=1, **{**ACDC, 'crosstalk': xt})
    g = gen(seed=2024, **kw)                                   # (no sample comes near 0 or 4095 here)
    assert g.kernel == 'grid'
    st = CubeStats(g)
    for i in range(E // chunk):
        st.accumulate(g.generate(chunk, first_event=i * chunk))
    r = st.result()
    n = E * P
    mean = r['bin_sum'] / n
    var = r['bin_sumsq'] / n - mean ** 2
    cfg = orc.make_config(**kw)
    B, n_drop, dt = 50, 10, 4.0
    h = np.asarray(cfg.pulse_template(np.arange(23) * dt), dtype=np.float64)
    lam = cfg.nsb_rate[0] * dt
    m1 = 1 / (1 - xt)
    m2 = xt / (1 - xt) ** 3 + m1 ** 2 + cfg.sigma_1[0] ** 2 * m1
    want_mean, want_var = np.zeros(B), np.zeros(B)
    for b in range(B):
        taps = h[:min(23, b + n_drop + 1)]                     # pe only exist from simulated bin 0 on
        want_mean[b] = cfg.baseline[0] + cfg.gain[0] * lam * m1 * taps.sum()
        want_var[b] = cfg.gain[0] ** 2 * lam * m2 * (taps ** 2).sum() + cfg.sigma_e[0] ** 2 + 1 / 12
    z = (mean - want_mean) / np.sqrt(want_var / n)
    assert np.abs(z).max() < 5, z
    assert np.abs(var / want_var - 1).max() < 0.01, (var / want_var)


def test_iterator_prefetches_into_two_reused_pinned_batches():
    """for event in generator (reference produce_data.py:119-126): events come out of two pinned batch
    slots that are allocated once; the next batch is generated while the current one is consumed;
    the events are the ones the bulk call gives."""
    g = gen(n_pixels=33, nsb_rate=0.3, n_photon=4., seed=17, n_events=53, batch_events=10, **ACDC)
    bulk, t_bulk, q_bulk = cube(gen(n_pixels=33, nsb_rate=0.3, n_photon=4., seed=17, **ACDC), 53, truth=True)
    slots = set()
    n = 0
    for ev in g:
        assert np.array_equal(np.asarray(ev.adc_count).astype(np.int16), bulk[ev.count])
        assert np.array_equal(ev.cherenkov_time, t_bulk[ev.count]) and np.array_equal(ev.cherenkov_photon, q_bulk[ev.count])
        slots.add(g._host[0].ctypes.data)
        n += 1
    assert n == 53 and g.count == 52
    assert len(slots) == 2                                   # batches alternate between the two slots
    assert {b[0].data_ptr() for b in g._bufs} == slots       # ... which were allocated exactly once
    # without prefetching: one slot, same events
    g1 = gen(n_pixels=33, nsb_rate=0.3, n_photon=4., seed=17, n_events=25, batch_events=10, prefetch=False, **ACDC)
    assert all(np.array_equal(np.asarray(ev.adc_count).astype(np.int16), bulk[ev.count]) for ev in g1)
    assert g1._bufs[1] is None
    # a bulk host call in between waits for the batch in flight and does not disturb the iterator
    g2 = gen(n_pixels=33, nsb_rate=0.3, n_photon=4., seed=17, n_events=30, batch_events=10, **ACDC)
    next(g2)
    extra = g2.generate_host(5, first_event=40, truth=False)
    assert np.array_equal(extra.view(np.int16), bulk[40:45])
    assert all(np.array_equal(np.asarray(ev.adc_count).astype(np.int16), bulk[ev.count]) for ev in g2)
    g.close(); g1.close(); g2.close()


def test_iterator_batches_of_a_long_run_grow_from_32_to_128_mib():
    """Open-ended (or >= 20000-event) runs: batches of 258, 258, 516, 1032, 1035 ... C4 events -- the first
    event arrives after a 32 MiB batch, the steady state runs on 128 MiB batches (generation and copy
    overlap inside a batch); each slot is re-pinned once, at the full size.  Same events as the bulk call."""
    import torch
    kw = dict(n_pixels=1296, nsb_rate=0.04, n_photon=2., seed=23, **ACDC)
    g = gen(**kw)
    assert (g._first_batch, g.batch_events) == (258, 1035)
    assert gen(n_events=5000, **kw).batch_events == 258                     # short run: small batches only
    bulk = gen(**kw).generate(2400)
    torch.cuda.synchronize()
    starts = []
    for ev in g:
        if ev.count == g._batch_first:
            starts.append(ev.count)
        if ev.count % 7 == 0 or ev.count == g._batch_first or ev.count == g._batch_first + g._batch_len - 1:
            assert np.array_equal(np.asarray(ev.adc_count).astype(np.int16), bulk[ev.count].cpu().numpy()), ev.count
        if ev.count == 2399:
            break
    assert starts == [0, 258, 516, 1032, 2064]
    assert g._bufs[0][0].shape[0] == 1035 and g._bufs[1][0].shape[0] == 1035
    g.close()


def test_pe_list_attributes_of_the_iterator():
    toy = gen(n_pixels=20, nsb_rate=0.3, n_photon=3., n_events=3, seed=8, **ACDC)
    assert toy.nsb_time.shape == (20, 1) and not toy.mask.any()          # before the first event (:194-196)
    for ev in toy:
        t, a, m = ev.nsb_time, ev.nsb_photon, ev.mask
        assert t.shape == a.shape == m.shape and t.shape[0] == 20
        n = m.sum(axis=1)
        assert abs(n.mean() - 0.3 * 240) < 5 * np.sqrt(72 / 20)
        for p in range(20):
            tt = t[p, :n[p]]
            assert np.all(np.diff(tt) >= 0) and tt.min() >= -40 and tt.max() < 200   # sorted, inside the window
            assert np.all(a[p, n[p]:] == 0)
        assert abs(a[m.astype(bool)].mean() - 1 / (1 - 0.08)) < 0.1


def test_usage_pattern_of_nsb_cluster_script():
    """reference nsb_cluster_7.py:18-44: reads adc_count before the first event, calls next()
    directly, does float arithmetic on the returned traces, reads toy.gain."""
    toy = gen(n_pixels=21, nsb_rate=1.0, seed=3)
    baseline = np.zeros(21)
    baseline += np.mean(toy.adc_count, axis=-1)
    assert not baseline.any()
    for i in range(3):
        toy.next()
        trace = toy.adc_count
        trace = trace - baseline[:, None]     # (the script's own `trace - baseline[j]` does not broadcast)
        trace[trace < 0] = 0
        trigger_trace = np.sum(trace, axis=0)
        assert trigger_trace.shape == (50,) and trigger_trace.min() > 21 * 200
    assert toy.gain.shape == (21,) and abs(toy.gain[0] - 5.8 / 2) < 1e-12 and toy.count == 2
