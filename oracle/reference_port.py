"""CPU oracle for the digicamtoy trace-generation path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product package (``digicamtoy_b200``) never
does: it fails loudly when its CUDA library is missing instead of falling back to this code.

What this is: a NumPy/SciPy restatement of the per-event Monte Carlo of the reference class
``NTraceGenerator`` (reference ``digicamtoy/tracegenerator/__init__.py:38-381``), written as plain
functions over an explicit ``numpy.random.RandomState`` so that

* with the same MT19937 seed it consumes random numbers in exactly the order the reference does
  (``next`` :201-226) and therefore reproduces the reference's ADC counts bit for bit, and
* every intermediate the GPU replay kernel needs (photo-electron times, smeared amplitudes,
  electronic noise) is returned instead of being thrown away.

Parity pin: ``tests/golden/*.npz`` were produced by importing the *real* reference (with the
two-line constructor patch described in SURVEY.md appendix B) in the build container, script
``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py`` checks this port against them
sample for sample.  The stochastic path of the reference has no seeded test of its own, so the
golden vectors generated from the reference are the pin.

The arithmetic deliberately keeps the reference's cost profile (dense (pixel, pe, bin) spline
evaluation, scalar Python recursion for crosstalk) because the same code is what ``bench.py`` times
as the CPU baseline.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np
from scipy.interpolate import interp1d

BACKWARD_NS = 40  # reference :93  (artificial_backward_time)

_UTILS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), os.pardir,
                          'digicamtoy_b200')


# ----------------------------------------------------------------------------------------------
# constructor math  (reference :52-157)
# ----------------------------------------------------------------------------------------------
def _poly_drop(rate_hz, ascending):
    """Cubic voltage-drop polynomials, zero at/above 1.4 GHz (reference utils/analytical.py:9-37)."""
    rate_hz = np.asarray(rate_hz, dtype=float)
    c0, c1, c2, c3 = ascending
    y = ((c3 * rate_hz + c2) * rate_hz + c1) * rate_hz + c0
    return y * (rate_hz < 1.4e9)


GAIN_DROP = (1.0000, -4.16866e-10, 2.27832e-19, -6.05033e-29)   # analytical.py:10
XT_DROP = (1., -7.41234e-10, 4.60472e-19, -1.28336e-28)         # analytical.py:20
PDE_DROP = (1., -2.2929e-10, 9.59912e-20, -2.22249e-29)         # analytical.py:30


def _vector(value, n_pixels):
    """Scalar or length-1 input is repeated per pixel, longer input is used as is (:52-78)."""
    a = np.atleast_1d(value)
    if a.shape[0] > 1:
        return a
    return np.array([a[0]] * n_pixels)


def _pulse_table(value, n_pixels):
    """(P, K) table of per-pixel pulse parameters; ragged rows padded with 0 (:59-65, :80-91)."""
    if isinstance(value, (int, float)) and not isinstance(value, bool):
        return np.atleast_2d([[value]] * n_pixels)
    rows = [np.atleast_1d(r) for r in (value.tolist() if isinstance(value, np.ndarray) else value)]
    width = max(len(r) for r in rows)
    out = np.zeros((len(rows), width), dtype=float)
    for i, r in enumerate(rows):
        out[i, :len(r)] = np.where(np.isnan(np.asarray(r, dtype=float)), 0.0, r)
    return out


@dataclass
class OracleConfig:
    """Everything ``__init__`` derives; attribute names follow the reference (row a11)."""
    n_pixels: int
    time_start: float      # already shifted by -40 ns (:96)
    time_end: float
    time_sampling: float
    n_samples: int
    nsb_rate: np.ndarray
    crosstalk: np.ndarray
    gain: np.ndarray       # effective gain after the NSB model (:112-128)
    sigma_e: np.ndarray
    sigma_1: np.ndarray    # already divided by the effective gain (:136)
    baseline: np.ndarray
    n_photon: np.ndarray
    time_signal: np.ndarray
    jitter: np.ndarray
    poisson: bool
    sub_binning: int
    tau: float
    norm: float            # max of the template on the fine grid (:148-149)
    true_baseline: np.ndarray
    template_x: np.ndarray
    template_y: np.ndarray  # normalised knots
    pulse_template: object = field(repr=False, default=None)

    @property
    def sampling_bins(self):
        return np.arange(self.time_start, self.time_end, self.time_sampling)   # :197-198

    @property
    def n_drop(self):
        return int(BACKWARD_NS // self.time_sampling)                           # :375


def resolve_pulse_file(pulse_shape_file):
    name = pulse_shape_file.lstrip('/')
    if os.path.isabs(pulse_shape_file) and os.path.exists(pulse_shape_file):
        return pulse_shape_file
    return os.path.normpath(os.path.join(_UTILS_DIR, name))


def make_config(time_start=0, time_end=200, time_sampling=4, n_pixels=1296, nsb_rate=0.6,
                crosstalk=0.08, gain_nsb=True, n_photon=0, poisson=True, sigma_e=0.8,
                sigma_1=0.8, gain=5.8, baseline=200, time_signal=20, jitter=0,
                pulse_shape_file='utils/pulse_SST-1M_pixel_0.dat', sub_binning=0,
                voltage_drop=False, fix_voltage_drop_broadcast=True, **_ignored):
    """Restates the constructor (:40-169).  ``fix_voltage_drop_broadcast`` selects the intended
    per-pixel scaling of ``n_photon`` instead of the reference's accidental (P,P) broadcast (:117)."""
    rate_set = _vector(nsb_rate, n_pixels)
    xt_set = _vector(crosstalk, n_pixels)
    npe = _pulse_table(n_photon, n_pixels)
    s_e = _vector(sigma_e, n_pixels)
    s_1 = _vector(sigma_1, n_pixels)
    g_set = _vector(gain, n_pixels)
    base = _vector(baseline, n_pixels)
    t_sig = _pulse_table(time_signal, n_pixels)
    jit = _pulse_table(jitter, n_pixels)

    n_samples = (time_end - time_start) // time_sampling                 # :95
    t0 = time_start - BACKWARD_NS                                        # :96

    if voltage_drop:                                                     # :112-117
        hz = rate_set * 1e9
        rate = rate_set * _poly_drop(hz, PDE_DROP)
        xt = xt_set * _poly_drop(hz, XT_DROP)
        g = g_set * _poly_drop(hz, GAIN_DROP)
        scale = _poly_drop(hz, PDE_DROP)
        npe = npe * (scale[:, None] if fix_voltage_drop_broadcast else scale)
    elif gain_nsb:                                                       # :119-122
        g = g_set / (1. + rate_set / gain_nsb)
        xt, rate = xt_set, rate_set
    else:                                                                # :124-128
        g, rate, xt = g_set, rate_set, xt_set

    s_1 = s_1 / g                                                        # :136

    x, y = np.loadtxt(resolve_pulse_file(pulse_shape_file), unpack=True)  # :139
    raw = interp1d(x, y, kind='cubic', bounds_error=False, fill_value=0., assume_sorted=True)
    fine_t = np.arange(t0, time_end, time_sampling / 100)                # :144-146
    fine_y = raw(fine_t)
    norm = np.max(fine_y)
    y_norm = y / norm                                                    # :149
    trapz = getattr(np, 'trapezoid', None) or np.trapz
    tau = trapz(fine_y / norm, fine_t)                                   # :150-151
    template = interp1d(x, y_norm, kind='cubic', bounds_error=False, fill_value=0.,
                        assume_sorted=True)                              # :152-154
    true_baseline = base + g * rate * (1 / (1 - xt)) * tau               # :156-157

    return OracleConfig(n_pixels=n_pixels, time_start=t0, time_end=time_end,
                        time_sampling=time_sampling, n_samples=n_samples, nsb_rate=rate,
                        crosstalk=xt, gain=g, sigma_e=s_e, sigma_1=s_1, baseline=base,
                        n_photon=npe, time_signal=t_sig, jitter=jit, poisson=poisson,
                        sub_binning=sub_binning, tau=float(tau), norm=float(norm),
                        true_baseline=true_baseline, template_x=x, template_y=y_norm,
                        pulse_template=template)


# ----------------------------------------------------------------------------------------------
# one event  (reference :201-381); every np.random call is made in the reference's order
# ----------------------------------------------------------------------------------------------
@dataclass
class EventRecord:
    nsb_time: np.ndarray        # (P, M) f64
    nsb_amplitude: np.ndarray   # (P, M) f64 = float32 smeared cluster size x mask (what :354 multiplies)
    sig_time: np.ndarray        # (P, K) f64
    sig_amplitude: np.ndarray   # (P, K) float32 after crosstalk + smearing (== cherenkov_photon)
    e_noise: np.ndarray         # (P, B') f64, all simulated bins incl. the dropped ones
    analog: np.ndarray          # (P, B') f64 before noise / baseline
    adc_float: np.ndarray       # (P, B) f64 rounded, before the unsigned cast
    nsb_cluster: np.ndarray     # (P, M) int cluster sizes before smearing (diagnostics)

    @property
    def adc_count(self):
        """What the reference exposes: ``round().astype(np.uint)`` (:380-381), wrap included."""
        return self.adc_float.astype(np.int64).astype(np.uint64)


def _galton_watson_extra(rng, xt):
    """Total progeny minus one of a Poisson(xt) branching process (:269-277), same draw order."""
    n = rng.poisson(xt)
    total = n
    for _ in range(n):
        total += _galton_watson_extra(rng, xt)
    return total


def next_event(cfg: OracleConfig, rng, reference_aliasing=False) -> EventRecord:
    """One event.  ``reference_aliasing`` reproduces a reference defect: with ``poisson=False``
    ``cherenkov_photon`` IS ``self.n_photon`` (:236-237) and the in-place ``+=`` of :327 makes the
    configured pulse size grow from event to event.  The build does not do that."""
    P = cfg.n_pixels
    bins = cfg.sampling_bins
    span = cfg.time_end - cfg.time_start

    nsb_time = np.zeros((P, 1))
    nsb_n = np.zeros((P, 1))
    mask = np.zeros((P, 1))
    sig_time = np.zeros(cfg.n_photon.shape)
    sig_n = np.zeros(cfg.n_photon.shape, dtype=int)

    # --- NSB primaries (:240-267)
    if np.any(cfg.nsb_rate > 0):
        if cfg.sub_binning <= 0:
            counts = rng.poisson(lam=span * cfg.nsb_rate)
            width = np.max(counts)
            nsb_time = rng.uniform(size=(P, width)) * span + cfg.time_start
            nsb_n = np.ones(nsb_time.shape, dtype=int)
            mask = (np.arange(width)[None, :] < counts[:, None]).astype(int)
        else:
            nsb_time = np.tile(bins, (P, 1))
            lam = np.tile(cfg.nsb_rate * cfg.time_sampling, (len(bins), 1)).T
            nsb_n = rng.poisson(lam=lam)
            mask = nsb_n > 0

    # --- signal pulses (:228-238)
    if np.any(cfg.n_photon > 0):
        width = np.array(cfg.jitter, dtype=float, copy=True)
        off = width <= 0
        width[off] = 1
        shift = rng.uniform(-width / 2, width / 2)
        shift[off] = 0
        sig_time = cfg.time_signal + shift
        sig_n = rng.poisson(lam=cfg.n_photon) if cfg.poisson else np.array(cfg.n_photon, copy=True)
        sig_n[cfg.n_photon <= 0] = 0

    # --- optical crosstalk (:297-327): pixel-major, NSB slots first, then the pulses
    if np.any(cfg.crosstalk > 0):
        nsb_extra = np.zeros(nsb_n.shape, dtype=int)
        sig_extra = np.zeros(sig_n.shape, dtype=int)
        for p in range(P):
            xt = cfg.crosstalk[p]
            for i in range(nsb_n.shape[-1]):
                if cfg.sub_binning > 0:
                    for _ in range(nsb_n[p, i]):
                        nsb_extra[p, i] += _galton_watson_extra(rng, xt)
                else:
                    nsb_extra[p, i] = _galton_watson_extra(rng, xt)
            for k in range(sig_n.shape[-1]):
                for _ in range(int(sig_n[p, k])):
                    sig_extra[p, k] += _galton_watson_extra(rng, xt)
        nsb_n = nsb_n + nsb_extra
        sig_n = sig_n + sig_extra
        if reference_aliasing and not cfg.poisson and np.any(cfg.n_photon > 0):
            cfg.n_photon += sig_extra.astype(cfg.n_photon.dtype)
    nsb_cluster = np.array(nsb_n, copy=True)

    # --- gain smearing (:329-348); amplitudes end up float32
    if np.any(cfg.sigma_1 > 0):
        s_nsb = np.sqrt(nsb_n) * cfg.sigma_1[:, None]
        s_sig = np.sqrt(sig_n) * cfg.sigma_1[:, None]
        z = s_nsb <= 0
        s_nsb[z] = 1
        d_nsb = rng.normal(0, s_nsb)
        d_nsb[z] = 0
        z = s_sig <= 0
        s_sig[z] = 1
        d_sig = rng.normal(0, s_sig)
        d_sig[z] = 0
        nsb_n = nsb_n.astype(np.float32)
        sig_n = sig_n.astype(np.float32)
        nsb_n += d_nsb
        sig_n += d_sig

    # --- shaping (:350-362): dense (P, M, B') evaluation, summed over the pe axis
    nsb_amp = nsb_n * mask
    analog = np.sum(cfg.pulse_template(bins - nsb_time[..., None]) * nsb_amp[..., None], axis=1)
    analog = analog + np.sum(cfg.pulse_template(bins - sig_time[..., None]) * sig_n[..., None],
                             axis=1)
    analog = analog * cfg.gain[:, None]

    # --- electronic noise (:364-371)
    e_noise = np.zeros(analog.shape)
    v = analog
    if np.any(cfg.sigma_e > 0):
        spread = np.tile(np.sqrt(cfg.sigma_e ** 2), (bins.shape[-1], 1)).T
        e_noise = rng.normal(0., spread)
        v = analog + e_noise

    # --- digitisation (:373-381)
    kept = v[:, cfg.n_drop:]
    adc_float = (kept + cfg.baseline[:, None]).round()

    return EventRecord(nsb_time=nsb_time, nsb_amplitude=np.asarray(nsb_amp, dtype=float),
                       sig_time=np.asarray(sig_time, dtype=float),
                       sig_amplitude=np.asarray(sig_n), e_noise=e_noise, analog=analog,
                       adc_float=adc_float, nsb_cluster=nsb_cluster)


# ----------------------------------------------------------------------------------------------
# deterministic back half on its own: what the replay kernel must reproduce
# ----------------------------------------------------------------------------------------------
def shape_and_digitise(cfg: OracleConfig, nsb_time, nsb_amp, sig_time, sig_amp, e_noise_kept):
    """Reference arithmetic of :350-381 applied to given pe lists; returns rounded f64 (P, B).

    ``e_noise_kept`` is the noise of the kept bins only, (P, B)."""
    bins = cfg.sampling_bins
    a = np.sum(cfg.pulse_template(bins - nsb_time[..., None]) * nsb_amp[..., None], axis=1)
    a = a + np.sum(cfg.pulse_template(bins - sig_time[..., None]) * sig_amp[..., None], axis=1)
    a = a * cfg.gain[:, None]
    kept = a[:, cfg.n_drop:] + e_noise_kept
    return (kept + cfg.baseline[:, None]).round()


def saturate(adc_float, lo=0, hi=4095):
    """The build's 12-bit clamp (the reference has none)."""
    return np.clip(adc_float, lo, hi).astype(np.int16)


# ----------------------------------------------------------------------------------------------
# closed-form per-bin moments of the stochastic path
# ----------------------------------------------------------------------------------------------
def closed_form_moments(cfg: OracleConfig, n_grid=4001):
    """Mean and variance of every kept ADC sample, (P, B) each, ignoring saturation.

    NSB is a compound Poisson process: cluster size n ~ Borel(xt) (E n = 1/(1-xt),
    Var n = xt/(1-xt)^3), amplitude n + sqrt(n) sigma_1 z.  NSB only exists after time_start, so
    early bins integrate a truncated template (the edge effect the GPU must reproduce)."""
    bins = cfg.sampling_bins[cfg.n_drop:]
    x_lo, x_hi = cfg.template_x[0], cfg.template_x[-1]
    trapz = getattr(np, 'trapezoid', None) or np.trapz
    I1 = np.zeros(len(bins))
    I2 = np.zeros(len(bins))
    for b, tb in enumerate(bins):
        # pe at time t in [time_start, time_end) contributes h(tb - t): x = tb - t
        hi = min(x_hi, tb - cfg.time_start)
        lo = max(x_lo, tb - cfg.time_end)
        if hi > lo:
            x = np.linspace(lo, hi, n_grid)
            h = cfg.pulse_template(x)
            I1[b] = trapz(h, x)
            I2[b] = trapz(h * h, x)
    xt = cfg.crosstalk[:, None]
    g = cfg.gain[:, None]
    r = cfg.nsb_rate[:, None]
    s1 = cfg.sigma_1[:, None]
    m1 = 1 / (1 - xt)
    m2 = xt / (1 - xt) ** 3 + m1 ** 2 + s1 ** 2 * m1
    mean = cfg.baseline[:, None] + g * r * m1 * I1[None, :]
    var = g ** 2 * r * m2 * I2[None, :] + cfg.sigma_e[:, None] ** 2 + 1 / 12

    # signal pulses: total pe S ~ generalised Poisson, time uniform within the jitter window
    K = cfg.n_photon.shape[1]
    for k in range(K):
        npe = np.clip(cfg.n_photon[:, k], 0, None)[:, None]
        if cfg.poisson:
            eS = npe * m1
            vS = npe / (1 - xt) ** 3
        else:
            npe = np.round(npe)
            eS = npe * m1
            vS = npe * xt / (1 - xt) ** 3
        eA, vA = eS, vS + s1 ** 2 * eS
        eh = np.zeros((cfg.n_pixels, len(bins)))
        eh2 = np.zeros((cfg.n_pixels, len(bins)))
        for p in range(cfg.n_pixels):
            j = cfg.jitter[p, k]
            ts = cfg.time_signal[p, k]
            if j > 0:
                tt = ts + np.linspace(-j / 2, j / 2, 201)
                hv = cfg.pulse_template(bins[None, :] - tt[:, None])
                eh[p] = trapz(hv, tt, axis=0) / j
                eh2[p] = trapz(hv * hv, tt, axis=0) / j
            else:
                hv = cfg.pulse_template(bins - ts)
                eh[p], eh2[p] = hv, hv * hv
        mean = mean + g * eA * eh
        var = var + g ** 2 * ((vA + eA ** 2) * eh2 - (eA * eh) ** 2)
    return mean, var


def borel_pmf(xt, n_max=40):
    """P(n), n = 1..n_max (reference utils/pdf.py:26-32 closed form)."""
    from scipy.special import gammaln
    n = np.arange(1, n_max + 1)
    if xt <= 0:
        return (n == 1).astype(float)
    return np.exp(-xt * n + (n - 1) * np.log(xt * n) - gammaln(n + 1))


# ----------------------------------------------------------------------------------------------
# convenience wrapper used by the CPU-baseline timing and the statistical tests
# ----------------------------------------------------------------------------------------------
class OracleGenerator:
    """Iterator facade over :func:`next_event` with the *intended* ``n_events`` semantics
    (exactly n events, reference test ``test_NTraceGenerator.py:4-13``)."""

    def __init__(self, seed=None, n_events=None, **kwargs):
        self.cfg = make_config(**kwargs)
        self.rng = np.random.RandomState(seed)
        self.n_events = n_events
        self.count = -1
        self.record = None

    def __iter__(self):
        return self

    def __next__(self):
        if self.n_events is not None and self.count + 1 >= self.n_events:
            raise StopIteration
        self.count += 1
        self.record = next_event(self.cfg, self.rng)
        return self

    @property
    def adc_count(self):
        return self.record.adc_count


def time_events(n_events, seed=0, **kwargs):
    """Seconds per event of the oracle on one core (CPU-baseline leg of bench.py)."""
    import time
    gen = OracleGenerator(seed=seed, **kwargs)
    next(gen)  # warm-up (imports, first-touch)
    t = time.perf_counter()
    for _ in range(n_events):
        next(gen)
    return (time.perf_counter() - t) / max(n_events, 1)
