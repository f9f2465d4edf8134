"""K1t, the tensor-core generation kernel (tcgen05.mma on per-cell pe moments): same photo-electron
stream as the sweep kernel, pulse shaping summed through fp16 moments x (hi + lo) fp16 coefficients
with fp32 accumulation.  Checked against the sweep kernel on the same seed, against the closed form,
and for independence of the launch geometry."""
import numpy as np
import pytest

from oracle import reference_port as orc

pytestmark = pytest.mark.gpu

ACDC = dict(time_start=0, time_end=200, time_sampling=4, crosstalk=0.08, gain_nsb=True, poisson=True,
            sigma_e=1.25, sigma_1=0.8, gain=5, baseline=280, time_signal=20, jitter=0.2)


def gen(**kw):
    from digicamtoy_b200 import NTraceGenerator
    return NTraceGenerator(**kw)


@pytest.mark.parametrize('rate,npe,P', [(1.0, 100., 300), (0.125, 10., 131), (3.0, 0., 64), (0.01, 2., 200)])
def test_tensor_kernel_matches_sweep_kernel_on_the_same_seed(rate, npe, P):
    import torch
    kw = dict(n_pixels=P, nsb_rate=rate, n_photon=npe, **ACDC)
    gs, gt = gen(seed=7, kernel='sweep', **kw), gen(seed=7, kernel='tensor', **kw)
    assert gs.kernel == 'sweep' and gt.kernel == 'tensor'
    n = 40
    adc_s, an_s = gs.generate_analog(n)
    adc_t, an_t = gt.generate_analog(n)
    torch.cuda.synchronize()
    gain_eff = float(gs.gain[0])
    d = (an_t.double() - an_s.double()).cpu().numpy() * gain_eff          # LSB
    # fp16 rounding of the moments: random, unbiased, ~1e-3 LSB per pe
    scale = max(1.0, np.sqrt(rate))
    assert np.sqrt((d ** 2).mean()) < 5e-3 * scale, np.sqrt((d ** 2).mean())
    assert abs(d.mean()) < 2e-4 * scale, d.mean()
    # the signal pulses do not go through fp16: a 100 pe pulse leaves the difference where it was
    dd = (adc_t.int() - adc_s.int()).cpu().numpy()
    assert np.abs(dd).max() <= 1
    assert (dd != 0).mean() < 5e-3 * scale


def test_tensor_kernel_is_independent_of_chunking_and_offsets():
    g = gen(n_pixels=37, nsb_rate=0.8, n_photon=8., seed=99, kernel='tensor', **ACDC)
    import torch
    def cube(n, first):
        out = g.generate(n, first_event=first)
        torch.cuda.synchronize()
        return out.cpu().numpy()
    whole = cube(40, 0)
    parts = np.concatenate([cube(7, 0), cube(13, 7), cube(20, 20)])
    assert np.array_equal(whole, parts)
    assert np.array_equal(cube(5, 17), whole[17:22])
    g2 = gen(n_pixels=37, nsb_rate=0.8, n_photon=8., seed=99, kernel='tensor', **ACDC)
    out = g2.generate(40, first_event=0); torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), whole)


@pytest.mark.parametrize('rate', [0.04, 1.0])
def test_tensor_kernel_moments_against_closed_form(rate):
    import torch
    kw = dict(n_pixels=64, nsb_rate=rate, n_photon=0., **ACDC)
    g = gen(seed=77, kernel='tensor', saturate=None, **kw)
    n_events = 3000
    c = g.generate(n_events); torch.cuda.synchronize()
    c = c.cpu().numpy().astype(np.float64)
    mean, var = orc.closed_form_moments(orc.make_config(**kw))
    z = (c.mean(axis=0) - mean) / np.sqrt(var / n_events)
    assert np.abs(z).max() < 6 and abs(z.mean()) < 6 / np.sqrt(z.size)
    ratio = c.var(axis=0).mean(axis=0) / var.mean(axis=0)
    assert np.abs(ratio - 1).max() < 0.03, ratio


def test_tensor_kernel_handles_per_pixel_rates_dark_pixels_and_partial_blocks():
    """Pixels without NSB, rates from 1 MHz to 2 GHz in one camera, a trace count that is not a
    multiple of the 128-trace CTA; windows the accumulator cannot hold are refused."""
    import torch
    rates = np.concatenate([np.zeros(3), np.logspace(-3, 0.3, 58)])
    kw = dict(n_pixels=61, nsb_rate=rates, n_photon=3., **ACDC)
    gs, gt = gen(seed=3, kernel='sweep', **kw), gen(seed=3, kernel='tensor', **kw)
    a, b = gs.generate(33), gt.generate(33)
    torch.cuda.synchronize()
    d = (a.int() - b.int()).cpu().numpy()
    assert np.abs(d).max() <= 1 and (d != 0).mean() < 5e-3
    assert np.array_equal(a[:, :3].cpu().numpy(), b[:, :3].cpu().numpy())      # no NSB: nothing goes through fp16
    with pytest.raises(ValueError):                          # more bins than the accumulator has columns for
        gen(seed=1, kernel='tensor', n_pixels=4, nsb_rate=1.0, **{**ACDC, 'time_end': 408})
    with pytest.raises(ValueError):                          # odd number of bins
        gen(seed=1, kernel='tensor', n_pixels=4, nsb_rate=1.0, **{**ACDC, 'time_end': 204})


@pytest.mark.parametrize('time_end', [368, 400])
def test_tensor_kernel_on_dark_yaml_length_windows(time_end):
    """92 bins (dark.yml, the commissioning grid of config_files/generate.py) and 100 bins: the kept
    bins start at accumulator column 10 and the last cells issue N = 16 MMAs so that nothing is
    written past column 127; every kept bin still gets all its taps."""
    import torch
    kw = dict(n_pixels=140, nsb_rate=1.2, n_photon=30., **{**ACDC, 'time_end': time_end, 'time_signal': 300.})
    gs, gt = gen(seed=4, kernel='sweep', **kw), gen(seed=4, kernel='tensor', **kw)
    assert gen(seed=4, **kw).kernel == 'tensor'              # AUTO takes it too
    (a, an_s), (b, an_t) = gs.generate_analog(40), gt.generate_analog(40)
    torch.cuda.synchronize()
    assert a.shape[2] == time_end // 4
    d = ((an_t.double() - an_s.double()) * float(gs.gain[0])).cpu().numpy()
    rms_per_bin = np.sqrt((d ** 2).mean(axis=(0, 1)))
    assert rms_per_bin.max() < 4e-3, rms_per_bin             # no bin misses a tap (that would be O(1) LSB)
    dd = (a.int() - b.int()).cpu().numpy()
    assert np.abs(dd).max() <= 1 and (dd != 0).mean() < 5e-3
    kwq = dict(kw, n_photon=0.)
    c = gen(seed=9, kernel='tensor', saturate=None, **kwq).generate(1500)
    torch.cuda.synchronize()
    c = c.cpu().numpy().astype(np.float64)
    mean, var = orc.closed_form_moments(orc.make_config(**kwq))
    z = (c.mean(axis=0) - mean) / np.sqrt(var / c.shape[0])
    assert np.abs(z).max() < 6


@pytest.mark.parametrize('window', [dict(time_start=0, time_end=320), dict(time_start=0, time_end=16),
                                    dict(time_start=-40, time_end=104), dict(time_start=0, time_end=8)])
def test_tensor_kernel_window_shapes(window):
    """The largest window the accumulator holds (80 bins), small ones (4 and 2 bins), one that does
    not start at 0: same photo-electrons as the sweep kernel, so the analog sums agree to the fp16
    rounding of the moments -- which scales with the template: the reference normalises it by its
    maximum INSIDE the simulated window (:144-149), 1.8x larger than usual for the 8 ns window."""
    import torch
    kw = dict(ACDC)
    kw.update(window)
    kw.update(n_pixels=70, nsb_rate=0.9, n_photon=20., time_signal=kw['time_start'] + 20)
    gs, gt = gen(seed=11, kernel='sweep', **kw), gen(seed=11, kernel='tensor', **kw)
    (a, an_s), (b, an_t) = gs.generate_analog(60), gt.generate_analog(60)
    torch.cuda.synchronize()
    assert a.shape[2] == (kw['time_end'] - kw['time_start']) // 4
    h_max = float(np.abs(gs.pulse_template(gs.params.knot_x)).max())
    assert 0.99 < h_max < 2.0
    rms = float(((an_t.double() - an_s.double()) * float(gs.gain[0])).pow(2).mean().sqrt())
    assert rms < 3.2e-3 * h_max, (rms, h_max)
    d = (a.int() - b.int()).cpu().numpy()
    assert np.abs(d).max() <= 1 and (d != 0).mean() < 3.5e-3 * h_max
    assert a.float().std() > 1.0


@pytest.mark.parametrize('window,P,E', [
    (dict(time_start=0, time_end=200), 1296, 80),     # 50 bins: five groups per SM, 810 items = more than one per group
    (dict(time_start=0, time_end=200), 37, 3),        # one partial item
    (dict(time_start=0, time_end=368), 1296, 70),     # 92 bins: four groups of 128 columns, six slots
    (dict(time_start=0, time_end=320), 131, 40),      # 80 bins
    (dict(time_start=0, time_end=16), 300, 33),       # 4 bins
    (dict(time_start=-40, time_end=104), 70, 60),     # window that does not start at 0
])
def test_persistent_tensor_kernel_equals_the_cta_kernel(monkeypatch, window, P, E):
    """k_generate_tensor_mega (one CTA per SM, 4-5 independent groups, each walking through its own
    128-trace items) does the arithmetic of k_generate_tensor in the same order: identical cubes,
    signal charges and analog sums, whatever the number of items per group."""
    import torch
    kw = dict(ACDC)
    kw.update(window)
    kw.update(n_pixels=P, nsb_rate=1.1, n_photon=12., time_signal=kw['time_start'] + 20)
    out = {}
    for mega in ('0', '1', '4'):
        monkeypatch.setenv('DCT_TENSOR_MEGA', mega)
        g = gen(seed=21, kernel='tensor', **kw)
        adc, analog = g.generate_analog(E)
        cube, t_sig, q_sig = g.generate(E, first_event=5, truth=True)
        torch.cuda.synchronize()
        out[mega] = (adc.clone(), analog.clone(), cube.clone(), t_sig.clone(), q_sig.clone())
        g.close()
    for mega in ('1', '4'):
        for x, y in zip(out['0'], out[mega]):
            assert torch.equal(x, y), mega
    assert out['0'][0].float().std() > 1.0


def test_auto_leaves_the_tensor_kernel_alone_when_its_rounding_would_show():
    """AUTO picks K1t only while the estimated fp16 rounding stays below 0.01 LSB: not for a gain of
    60 LSB per pe, nor for a window that starts after the pulse maximum (template scaled up ~30x)."""
    base = dict(n_pixels=8, nsb_rate=1.0, n_photon=0., **ACDC)
    assert gen(seed=1, **base).kernel == 'tensor'
    assert gen(seed=1, **{**base, 'gain': 60., 'gain_nsb': False}).kernel == 'sweep'
    assert gen(seed=1, **{**base, 'time_start': 100, 'time_end': 300}).kernel == 'sweep'
    assert gen(seed=1, kernel='tensor', **{**base, 'gain': 60., 'gain_nsb': False}).kernel == 'tensor'   # explicit choice stands


def test_two_tensor_generators_on_two_streams_share_the_tensor_memory():
    """Two K1t launches in flight at once: every SM's 512 TMEM columns are taken by whichever CTAs
    arrive first, the others wait in tcgen05.alloc -- results must not depend on it."""
    import torch
    kw = dict(n_pixels=1296, nsb_rate=1.0, n_photon=5., **ACDC)
    g1, g2 = gen(seed=1, kernel='tensor', **kw), gen(seed=2, kernel='tensor', **kw)
    ref1, ref2 = g1.generate(300, first_event=0), g2.generate(300, first_event=0)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(s1):
        o1 = g1.generate(300, first_event=0)
    with torch.cuda.stream(s2):
        o2 = g2.generate(300, first_event=0)
    torch.cuda.synchronize()
    assert torch.equal(o1, ref1) and torch.equal(o2, ref2)


def test_randomised_cross_checks_of_all_generation_kernels():
    """tests/tools/fuzz_kernels.py: 150 random configurations (1..300 pixels, 2..130 bins, zero to
    3 GHz and per-pixel rates, pulses of 0..3000 pe, crosstalk 0..0.3): sweep == scatter bit for bit,
    tensor within rounding of them, chunked == one-shot, fused statistics == separate == NumPy."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        'fuzz_kernels', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tools', 'fuzz_kernels.py'))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    rs = np.random.RandomState(2026)
    kernels_seen = set()
    for i in range(150):
        msg = fuzz.run_case(i, rs)
        kernels_seen.add(msg.rsplit('auto=', 1)[-1] if 'auto=' in msg else 'skipped')
    assert {'sweep', 'tensor', 'scatter'} <= kernels_seen | {'scatter'}
