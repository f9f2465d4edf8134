"""T1: replay parity.  The reference's own photo-electron lists and noise (golden fixtures, and
fresh oracle events at full camera size) go through the GPU shaping + digitisation via the C ABI;
ADC counts must equal the reference's wherever the reference value fits the 12-bit range."""
import numpy as np
import pytest

from oracle import reference_port as orc

from helpers import GOLDEN_CASES, load_golden, replay_on_gpu

pytestmark = pytest.mark.gpu


def _handle(kwargs, n_pulses):
    from digicamtoy_b200 import native
    from digicamtoy_b200.params import build_params
    kw = {k: v for k, v in kwargs.items()}
    p = build_params(**kw)
    if p.n_pulses != n_pulses:            # reference (P,P) broadcast case: only K matters for replay
        kw['n_photon'] = np.zeros((p.n_pixels, n_pulses))
        p2 = build_params(**kw)
        p2.gain, p2.baseline = p.gain, p.baseline
        p = p2
    return native.Handle(p, device=0), p


def _events_from_golden(z, meta, n_drop):
    evs, refs = [], []
    for e in range(meta['n_events']):
        st, sa = np.broadcast_arrays(z['sig_time_%d' % e], z['sig_amp_%d' % e])
        evs.append((z['nsb_time_%d' % e], z['nsb_amp_%d' % e], st, sa, z['e_noise_%d' % e][:, n_drop:]))
        refs.append(z['adc_%d' % e])
    return evs, np.stack(refs)


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_replay_f64_equals_reference(name):
    z, meta = load_golden(name)
    K = np.broadcast_arrays(z['sig_time_0'], z['sig_amp_0'])[0].shape[1]
    h, p = _handle(meta['kwargs'], K)
    evs, ref = _events_from_golden(z, meta, p.n_drop)
    adc, n_clamped = replay_on_gpu(h, 'f64', evs, p.n_pixels, p.n_bins)
    in_range = (ref >= 0) & (ref <= 4095)
    assert np.array_equal(adc[in_range], ref[in_range])          # bit-exact, zero mismatches
    assert np.array_equal(adc[~in_range], np.clip(ref[~in_range], 0, 4095))
    assert n_clamped == int((~in_range).sum())
    h.close()


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_replay_f32_within_one_lsb(name):
    z, meta = load_golden(name)
    K = np.broadcast_arrays(z['sig_time_0'], z['sig_amp_0'])[0].shape[1]
    h, p = _handle(meta['kwargs'], K)
    evs, ref = _events_from_golden(z, meta, p.n_drop)
    adc, _ = replay_on_gpu(h, 'f32', evs, p.n_pixels, p.n_bins)
    diff = np.abs(adc.astype(np.int64) - np.clip(ref, 0, 4095))
    assert diff.max() <= 1                                        # rounding ties only
    assert (diff != 0).mean() <= 2e-3, 'float32 mismatches: %d of %d' % ((diff != 0).sum(), diff.size)
    h.close()


def test_replay_full_camera_one_ghz():
    """C4-like event at the full 1296 pixels from the oracle (seeded), incl. saturating pixels."""
    P = 1296
    kw = dict(time_start=0, time_end=200, time_sampling=4, n_pixels=P, nsb_rate=1.0, crosstalk=0.08,
              gain_nsb=True, n_photon=[[float(10 ** (3.5 * i / (P - 1)))] for i in range(P)],
              poisson=True, sigma_e=1.25, sigma_1=0.8, gain=5, baseline=280, time_signal=20, jitter=0.2)
    cfg = orc.make_config(**kw)
    rec = orc.next_event(cfg, np.random.RandomState(2026))
    h, p = _handle(kw, 1)
    ev = [(rec.nsb_time, rec.nsb_amplitude, rec.sig_time, np.asarray(rec.sig_amplitude, dtype=float),
           rec.e_noise[:, cfg.n_drop:])]
    ref = rec.adc_float.astype(np.int64)
    in_range = (ref >= 0) & (ref <= 4095)
    assert (~in_range).sum() > 0                                  # saturation is exercised
    adc64, n_clamped = replay_on_gpu(h, 'f64', ev, P, p.n_bins)
    assert np.array_equal(adc64[0][in_range], ref[in_range])
    assert np.array_equal(adc64[0], np.clip(ref, 0, 4095))
    assert n_clamped == int((~in_range).sum())
    adc32, _ = replay_on_gpu(h, 'f32', ev, P, p.n_bins)
    diff = np.abs(adc32[0].astype(np.int64) - np.clip(ref, 0, 4095))
    n_bad = int((diff != 0).sum())
    print('float32 replay: %d of %d samples differ by 1 LSB' % (n_bad, diff.size))
    assert diff.max() <= 1 and n_bad <= 1e-3 * diff.size
    h.close()
