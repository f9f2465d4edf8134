"""Regenerates tests/golden/*.npz by importing the REAL reference.  Build-container only.

The reference (``/root/reference``, read-only, python) cannot run at HEAD: its constructor raises
at ``digicamtoy/tracegenerator/__init__.py:159-164`` and ``np.int`` (:193) no longer exists.  This
script copies the package to a temporary directory, applies exactly those two one-token fixes, and
runs ``NTraceGenerator.next()`` under fixed MT19937 seeds, recording for every event

  * the state the reference leaves behind after photon smearing (pe times / amplitudes, :217-218),
  * the electronic-noise array drawn at :369 (captured by wrapping ``np.random.normal``),
  * the final ADC counts (:381), stored as int64 (a wrapped negative shows up as < 0).

Nothing under tests/ or the product reads /root/reference at run time; only the committed .npz
fixtures travel.  Usage:  python tests/golden/make_golden.py
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

REFERENCE = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))

P0 = 'utils/pulse_SST-1M_pixel_0.dat'
LOWGAIN = 'utils/pulse_SST-1M_AfterPreampLowGain.dat'

CASES = {
    # name: (seed, n_events, kwargs)
    'acdc_nsb125_npe10': (11, 3, dict(time_start=0, time_end=200, time_sampling=4, n_pixels=12,
                                      nsb_rate=0.125, crosstalk=0.08, gain_nsb=True, n_photon=10.,
                                      poisson=True, sigma_e=1.25, sigma_1=0.8, gain=5, baseline=280,
                                      time_signal=20, jitter=0.2, pulse_shape_file=P0)),
    'dark_npe2': (12, 3, dict(time_start=0, time_end=368, time_sampling=4, n_pixels=12,
                              nsb_rate=0.003, crosstalk=0.08, gain_nsb=True, n_photon=2,
                              poisson=True, sigma_e=0.8, sigma_1=0.8, gain=5.8, baseline=200,
                              time_signal=0., jitter=0.3, pulse_shape_file=P0)),
    'c4_saturating': (13, 2, dict(time_start=0, time_end=200, time_sampling=4, n_pixels=8,
                                  nsb_rate=1.0, crosstalk=0.08, gain_nsb=True,
                                  n_photon=[[float(10 ** (3.5 * p / 7))] for p in range(8)],
                                  poisson=True, sigma_e=1.25, sigma_1=0.8, gain=5, baseline=280,
                                  time_signal=20, jitter=0.2, pulse_shape_file=P0)),
    'example_ragged': (14, 2, dict(pulse_shape_file=P0, n_pixels=3, baseline=[0, 500, 0],
                                   n_photon=[[100, 133], [60, 60, 60], [100, 133]],
                                   time_signal=[[0, 100], [20, 50, 100], [0, 100]],
                                   gain=[10, 10, 29], poisson=True)),
    'sub_binning': (15, 2, dict(n_pixels=6, nsb_rate=0.3, sub_binning=1, n_photon=4,
                                pulse_shape_file=P0)),
    'voltage_drop_refbug': (16, 2, dict(n_pixels=5, nsb_rate=0.4, voltage_drop=True, n_photon=20.,
                                        jitter=1.0, pulse_shape_file=P0)),
    'lowgain_template': (17, 2, dict(n_pixels=6, nsb_rate=0.5, gain_nsb=1 / 0.85, gain=5.8,
                                     baseline=500, sigma_e=0.8, sigma_1=0.8, crosstalk=0.08,
                                     time_end=200, pulse_shape_file=LOWGAIN)),
    'no_noise_no_xt': (18, 2, dict(n_pixels=4, nsb_rate=0, crosstalk=0, sigma_e=0, sigma_1=0,
                                   n_photon=5, gain_nsb=False, pulse_shape_file=P0)),
    'dt2_100ns': (19, 2, dict(time_start=0, time_end=100, time_sampling=2, n_pixels=4,
                              nsb_rate=0.2, pulse_shape_file=P0)),
    'per_pixel_arrays': (20, 2, dict(n_pixels=5, nsb_rate=[0.0, 0.05, 0.2, 0.6, 1.2],
                                     crosstalk=[0.0, 0.05, 0.1, 0.2, 0.3],
                                     sigma_e=[0.5, 0.8, 1.0, 1.5, 2.0],
                                     sigma_1=[0.0, 0.4, 0.8, 1.0, 1.2],
                                     gain=[4., 5., 5.8, 6., 7.], baseline=[100, 200, 300, 400, 10],
                                     n_photon=[[0.], [3.], [30.], [300.], [1.]],
                                     time_signal=[[10.], [20.], [30.], [40.], [-20.]],
                                     jitter=[[0.], [0.5], [1.], [2.], [4.]], gain_nsb=0.7,
                                     pulse_shape_file=P0)),
    'negative_wrap': (22, 2, dict(n_pixels=4, baseline=0, nsb_rate=0.01, sigma_e=2.0, n_photon=0,
                                  pulse_shape_file=P0)),
    'poisson_false': (21, 2, dict(n_pixels=4, nsb_rate=0.1, n_photon=7, poisson=False, crosstalk=0.1,
                                  pulse_shape_file=P0)),
}


def patched_reference():
    tmp = tempfile.mkdtemp(prefix='digicamtoy_ref_')
    shutil.copytree(os.path.join(REFERENCE, 'digicamtoy'), os.path.join(tmp, 'digicamtoy'))
    path = os.path.join(tmp, 'digicamtoy', 'tracegenerator', '__init__.py')
    src = open(path).read()
    a = ("                        self.true_baseline.mean()),\n"
         "              self.true_baseline.mean() - self.baseline.mean())")
    b = ("                        self.true_baseline.mean(),\n"
         "              self.true_baseline.mean() - self.baseline.mean()))")
    assert a in src and 'dtype=np.int)' in src
    src = src.replace(a, b).replace('dtype=np.int)', 'dtype=int)')
    open(path, 'w').write(src)
    return tmp


def main():
    tmp = patched_reference()
    sys.path.insert(0, tmp)
    import warnings
    warnings.simplefilter('ignore')
    from digicamtoy.tracegenerator import NTraceGenerator  # the reference class

    real_normal = np.random.normal
    normals = []

    def recording_normal(*a, **k):
        out = real_normal(*a, **k)
        normals.append(out)
        return out

    np.random.normal = recording_normal
    try:
        for name, (seed, n_events, kwargs) in CASES.items():
            np.random.seed(seed)
            gen = NTraceGenerator(**kwargs)
            out = {}
            for e in range(n_events):
                del normals[:]
                gen.next()
                amp = np.asarray(gen.nsb_photon * gen.mask, dtype=np.float64)
                out['nsb_time_%d' % e] = np.asarray(gen.nsb_time, dtype=np.float64)
                out['nsb_amp_%d' % e] = amp
                out['sig_time_%d' % e] = np.asarray(gen.cherenkov_time, dtype=np.float64)
                out['sig_amp_%d' % e] = np.asarray(gen.cherenkov_photon, dtype=np.float64)
                noise = normals[-1] if np.any(gen.sigma_e > 0) else np.zeros(
                    (gen.n_pixels, gen.adc_count.shape[1] + 40 // gen.time_sampling))
                out['e_noise_%d' % e] = np.asarray(noise, dtype=np.float64)
                out['adc_%d' % e] = gen.adc_count.astype(np.int64)   # uint64 wrap -> negative
            out['true_baseline'] = gen.true_baseline
            out['tau'] = np.float64(gen.tau)
            out['gain'] = gen.gain
            out['sigma_1'] = gen.sigma_1
            out['nsb_rate'] = gen.nsb_rate
            out['crosstalk'] = gen.crosstalk
            out['n_photon'] = np.asarray(gen.n_photon, dtype=np.float64)
            out['template_probe_x'] = np.linspace(-10, 300, 1241)
            out['template_probe_y'] = gen.pulse_template(out['template_probe_x'])
            out['meta'] = np.array(json.dumps(dict(seed=seed, n_events=n_events, kwargs=kwargs)))
            np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
            print(name, 'events', n_events, 'adc range',
                  min(out['adc_%d' % e].min() for e in range(n_events)),
                  max(out['adc_%d' % e].max() for e in range(n_events)))
    finally:
        np.random.normal = real_normal
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == '__main__':
    main()
