"""T0: the CPU oracle against vectors produced by the real reference (tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import reference_port as orc

from conftest import GOLDEN_DIR

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    meta = json.loads(str(z['meta']))
    return z, meta


def test_fixture_list():
    assert len(CASES) >= 12


@pytest.mark.parametrize('name', CASES)
def test_same_seed_reproduces_reference_bit_for_bit(name):
    """Same MT19937 seed, same draw order => identical pe lists, noise and ADC counts."""
    z, meta = load(name)
    kwargs = dict(meta['kwargs'])
    if name == 'voltage_drop_refbug':
        kwargs['fix_voltage_drop_broadcast'] = False   # reproduce the (P,P) broadcast of :117
    cfg = orc.make_config(**kwargs)
    np.testing.assert_array_equal(cfg.true_baseline, z['true_baseline'])
    assert cfg.tau == float(z['tau'])
    np.testing.assert_array_equal(cfg.gain, z['gain'])
    np.testing.assert_array_equal(cfg.sigma_1, z['sigma_1'])
    np.testing.assert_array_equal(cfg.pulse_template(z['template_probe_x']), z['template_probe_y'])
    rng = np.random.RandomState(meta['seed'])
    for e in range(meta['n_events']):
        rec = orc.next_event(cfg, rng, reference_aliasing=True)
        np.testing.assert_array_equal(rec.nsb_time, z['nsb_time_%d' % e])
        np.testing.assert_array_equal(rec.nsb_amplitude, z['nsb_amp_%d' % e])
        np.testing.assert_array_equal(rec.sig_time, z['sig_time_%d' % e])
        np.testing.assert_array_equal(np.asarray(rec.sig_amplitude, dtype=float), z['sig_amp_%d' % e])
        np.testing.assert_array_equal(rec.e_noise, z['e_noise_%d' % e])
        np.testing.assert_array_equal(rec.adc_count.astype(np.int64), z['adc_%d' % e])


@pytest.mark.parametrize('name', CASES)
def test_back_half_from_recorded_state(name):
    """shape_and_digitise (what the GPU replay must equal) from the reference's own state."""
    z, meta = load(name)
    kwargs = dict(meta['kwargs'])
    if name == 'voltage_drop_refbug':
        kwargs['fix_voltage_drop_broadcast'] = False
    cfg = orc.make_config(**kwargs)
    for e in range(meta['n_events']):
        noise = z['e_noise_%d' % e][:, cfg.n_drop:]
        adc = orc.shape_and_digitise(cfg, z['nsb_time_%d' % e], z['nsb_amp_%d' % e],
                                     z['sig_time_%d' % e], z['sig_amp_%d' % e], noise)
        ref = z['adc_%d' % e]
        # (analog + noise)[kept] == analog[kept] + noise[kept] exactly, so this is bit-exact too
        np.testing.assert_array_equal(adc.astype(np.int64), ref)


def test_negative_values_wrap_in_reference_and_clamp_in_build():
    z, meta = load('negative_wrap')
    adc = np.concatenate([z['adc_%d' % e].ravel() for e in range(meta['n_events'])])
    assert (adc < 0).any()            # uint64 wrap, viewed as int64
    assert orc.saturate(adc.astype(float)).min() == 0


def test_saturation_case_exceeds_12_bits():
    z, meta = load('c4_saturating')
    assert max(z['adc_%d' % e].max() for e in range(meta['n_events'])) > 4095


def test_intended_n_events():
    for n in (0, 1, 10):
        g = orc.OracleGenerator(seed=1, n_events=n, n_pixels=2, nsb_rate=1e-6)
        assert len([e for e in g]) == n


def test_borel_pmf_normalised_and_mean():
    p = orc.borel_pmf(0.08, 60)
    n = np.arange(1, 61)
    assert abs(p.sum() - 1) < 1e-12
    assert abs((p * n).sum() - 1 / (1 - 0.08)) < 1e-10


def test_closed_form_matches_oracle_sampling():
    kw = dict(n_pixels=64, nsb_rate=0.3, n_photon=0, sigma_e=1.0, gain=5.0, baseline=100,
              time_end=120)
    cfg = orc.make_config(**kw)
    rng = np.random.RandomState(5)
    acc = np.stack([orc.next_event(cfg, rng).adc_float for _ in range(40)])   # (E,P,B)
    mean, var = orc.closed_form_moments(cfg)
    n = acc.shape[0] * acc.shape[1]
    m_hat = acc.mean(axis=(0, 1))
    v_hat = acc.var(axis=(0, 1))
    assert np.all(np.abs(m_hat - mean[0]) < 5 * np.sqrt(var[0] / n))
    assert np.all(np.abs(v_hat / var[0] - 1) < 0.2)
