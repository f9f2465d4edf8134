"""Host-side parameter layer: everything ``NTraceGenerator.__init__`` derives, in float64.

Restates the constructor arithmetic of the reference
(``digicamtoy/tracegenerator/__init__.py:52-157``) and its voltage-drop polynomials
(``digicamtoy/utils/analytical.py:9-37``) and produces the flat arrays ``dct_create`` uploads.
Nothing here touches the GPU; the per-event Monte Carlo lives in ``csrc/``.
"""
from __future__ import annotations

import numbers
import os
from dataclasses import dataclass

import numpy as np

ARTIFICIAL_BACKWARD_TIME = 40          # ns simulated before time_start (reference :93)
PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))

# cubic polynomials in the NSB rate [Hz], ascending powers; zero at and above 1.4 GHz
_GAIN_DROP = (1.0000, -4.16866e-10, 2.27832e-19, -6.05033e-29)     # analytical.py:10
_XT_DROP = (1., -7.41234e-10, 4.60472e-19, -1.28336e-28)           # analytical.py:20
_PDE_DROP = (1., -2.2929e-10, 9.59912e-20, -2.22249e-29)           # analytical.py:30


def _drop(coefficients, rate_hz):
    rate_hz = np.asarray(rate_hz, dtype=np.float64)
    value = np.polyval(coefficients[::-1], rate_hz)
    return value * (rate_hz < 1.4e9)


def gain_drop(nsb_rate, cell_capacitance=85. * 1E-15, bias_resistance=10. * 1E3):
    """Simple RC gain-drop model (reference utils/analytical.py:4-6); not used by the generator."""
    return 1. / (1. + np.asarray(nsb_rate, dtype=np.float64) * cell_capacitance * bias_resistance * 1E9)


def true_gain_drop(nsb_rate_hz):
    return _drop(_GAIN_DROP, nsb_rate_hz)


def true_xt_drop(nsb_rate_hz):
    return _drop(_XT_DROP, nsb_rate_hz)


def true_pde_drop(nsb_rate_hz):
    return _drop(_PDE_DROP, nsb_rate_hz)


# ------------------------------------------------------------------------------------------------
# pulse template: not-a-knot cubic spline == scipy interp1d(kind='cubic')  (reference :141, :152)
# ------------------------------------------------------------------------------------------------
def not_a_knot_cubic(x, y):
    """Piecewise-cubic coefficients ``c[i] = (c0, c1, c2, c3)`` with
    ``S(t) = c0 + c1 u + c2 u^2 + c3 u^3``, ``u = t - x[i]`` on ``[x[i], x[i+1]]``."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.size
    if n < 4 or y.size != n:
        raise ValueError('pulse template needs >= 4 knots and matching amplitudes')
    h = np.diff(x)
    if np.any(h <= 0):
        raise ValueError('pulse template times must be strictly increasing')
    d = np.diff(y) / h
    A = np.zeros((n, n))
    rhs = np.zeros(n)
    for i in range(1, n - 1):
        A[i, i - 1] = h[i]
        A[i, i] = 2.0 * (h[i - 1] + h[i])
        A[i, i + 1] = h[i - 1]
        rhs[i] = 3.0 * (h[i] * d[i - 1] + h[i - 1] * d[i])
    w = x[2] - x[0]
    A[0, 0], A[0, 1] = h[1], w
    rhs[0] = ((h[0] + 2.0 * w) * h[1] * d[0] + h[0] ** 2 * d[1]) / w
    w = x[-1] - x[-3]
    A[-1, -1], A[-1, -2] = h[-2], w
    rhs[-1] = (h[-1] ** 2 * d[-2] + (2.0 * w + h[-1]) * h[-2] * d[-1]) / w
    s = np.linalg.solve(A, rhs)
    t = (s[:-1] + s[1:] - 2.0 * d) / h
    coef = np.empty((n - 1, 4))
    coef[:, 0] = y[:-1]
    coef[:, 1] = s[:-1]
    coef[:, 2] = (d - s[:-1]) / h - t
    coef[:, 3] = t / h
    return coef


def evaluate_template(knot_x, coef, t):
    """Spline value, 0 outside ``[knot_x[0], knot_x[-1]]`` (inclusive ends, fill_value=0)."""
    t = np.asarray(t, dtype=np.float64)
    idx = np.clip(np.searchsorted(knot_x, t, side='right') - 1, 0, len(knot_x) - 2)
    u = t - knot_x[idx]
    c = coef[idx]
    value = ((c[..., 3] * u + c[..., 2]) * u + c[..., 1]) * u + c[..., 0]
    return np.where((t >= knot_x[0]) & (t <= knot_x[-1]), value, 0.0)


def resolve_pulse_shape_file(pulse_shape_file):
    """The reference joins the name onto the package directory (:105-106); its own default has a
    leading '/' that breaks that join, so a leading '/' is dropped unless the path really exists."""
    if os.path.isabs(pulse_shape_file) and os.path.isfile(pulse_shape_file):
        return pulse_shape_file
    candidate = os.path.join(PACKAGE_DIR, pulse_shape_file.lstrip('/'))
    if not os.path.isfile(candidate):
        raise FileNotFoundError('pulse shape file {!r} not found (looked at {})'.format(
            pulse_shape_file, candidate))
    return candidate


# ------------------------------------------------------------------------------------------------
# broadcasting rules of the constructor
# ------------------------------------------------------------------------------------------------
def _is_scalar(v):
    return isinstance(v, numbers.Real) and not isinstance(v, bool) or (
        isinstance(v, np.ndarray) and v.ndim == 0)


def _per_pixel(name, value, n_pixels):
    a = np.atleast_1d(np.asarray(value, dtype=np.float64))
    if a.ndim != 1:
        raise ValueError('{} must be a scalar or a 1-d sequence'.format(name))
    if a.shape[0] == 1:
        return np.full(n_pixels, a[0])
    if a.shape[0] != n_pixels:
        raise ValueError('{} has {} entries for {} pixels'.format(name, a.shape[0], n_pixels))
    return a.copy()


def _per_pulse(name, value, n_pixels):
    """(P, K) table; ragged rows are padded with 0 like ``pd.DataFrame(x).fillna(0)`` (:65,:85,:91)."""
    if _is_scalar(value):
        return np.full((n_pixels, 1), float(value))
    rows = [np.atleast_1d(np.asarray(r, dtype=np.float64)) for r in value]
    if len(rows) != n_pixels:
        raise ValueError('{} has {} rows for {} pixels'.format(name, len(rows), n_pixels))
    width = max(r.shape[0] for r in rows)
    out = np.zeros((n_pixels, width))
    for i, r in enumerate(rows):
        out[i, :r.shape[0]] = np.nan_to_num(r, nan=0.0)
    return out


@dataclass
class TraceParams:
    n_pixels: int
    n_pulses: int
    n_samples: int            # reference attribute (:95): (time_end - time_start) // time_sampling
    n_bins: int               # bins actually produced: len(arange(t0, time_end, dt)) - n_drop
    n_drop: int
    time_start: float         # shifted by -40 ns, like the reference attribute (:96)
    time_end: float
    time_sampling: float
    nsb_rate: np.ndarray
    crosstalk: np.ndarray
    gain: np.ndarray
    sigma_e: np.ndarray
    sigma_1: np.ndarray
    baseline: np.ndarray
    n_photon: np.ndarray
    time_signal: np.ndarray
    jitter: np.ndarray
    poisson: bool
    sub_binning: int
    knot_x: np.ndarray
    coef: np.ndarray          # (n_knots-1, 4) of the NORMALISED template
    norm: float
    tau: float
    true_baseline: np.ndarray
    set_nsb_rate: np.ndarray
    set_crosstalk: np.ndarray
    set_gain: np.ndarray

    @property
    def sampling_bins(self):
        """Times of the kept bins (what ``sampling_bins`` holds after ``next``, :378)."""
        return np.arange(self.time_start, self.time_end, self.time_sampling)[self.n_drop:]


def build_params(time_start=0, time_end=200, time_sampling=4, n_pixels=1296, nsb_rate=0.6,
                 crosstalk=0.08, gain_nsb=True, n_photon=0, poisson=True, sigma_e=0.8, sigma_1=0.8,
                 gain=5.8, baseline=200, time_signal=20, jitter=0,
                 pulse_shape_file='/utils/pulse_SST-1M_pixel_0.dat', sub_binning=0,
                 voltage_drop=False):
    n_pixels = int(n_pixels)
    if n_pixels <= 0:
        raise ValueError('n_pixels must be positive')
    if not time_sampling > 0 or not time_end > time_start:
        raise ValueError('need time_sampling > 0 and time_end > time_start')

    rate_set = _per_pixel('nsb_rate', nsb_rate, n_pixels)
    xt_set = _per_pixel('crosstalk', crosstalk, n_pixels)
    gain_set = _per_pixel('gain', gain, n_pixels)
    s_e = _per_pixel('sigma_e', sigma_e, n_pixels)
    s_1 = _per_pixel('sigma_1', sigma_1, n_pixels)
    base = _per_pixel('baseline', baseline, n_pixels)
    npe = _per_pulse('n_photon', n_photon, n_pixels)
    t_sig = _per_pulse('time_signal', time_signal, n_pixels)
    jit = _per_pulse('jitter', jitter, n_pixels)
    try:
        npe, t_sig, jit = (np.ascontiguousarray(a) for a in np.broadcast_arrays(npe, t_sig, jit))
    except ValueError:
        raise ValueError('n_photon {}, time_signal {} and jitter {} do not broadcast to one '
                         '(pixel, pulse) table'.format(npe.shape, t_sig.shape, jit.shape))

    if np.any(rate_set < 0):
        raise ValueError('nsb_rate must be >= 0')

    n_samples = int((time_end - time_start) // time_sampling)            # :95
    t0 = time_start - ARTIFICIAL_BACKWARD_TIME                           # :96

    if voltage_drop:                                                     # :112-117
        hz = rate_set * 1e9
        pde = true_pde_drop(hz)
        rate = rate_set * pde
        xt = xt_set * true_xt_drop(hz)
        g = gain_set * true_gain_drop(hz)
        npe = npe * pde[:, None]        # per pixel; the reference broadcasts (P,1)*(P,) by accident
    elif gain_nsb:                                                       # :119-122 (True == 1.0 GHz)
        g = gain_set / (1.0 + rate_set / float(gain_nsb))
        rate, xt = rate_set, xt_set
    else:                                                                # :124-128
        g, rate, xt = gain_set, rate_set, xt_set
    if np.any((xt < 0) | (xt >= 1)):
        raise ValueError('crosstalk must lie in [0, 1)')
    if np.any(g == 0) and np.any(s_1 != 0):
        raise ValueError('effective gain is 0 for some pixel; sigma_1 / gain is undefined')
    with np.errstate(divide='ignore', invalid='ignore'):
        s_1_rel = np.where(g != 0, s_1 / g, 0.0)                         # :136

    # template: interpolate, normalise by the maximum on the dt/100 grid, integrate (:139-154)
    knot_x, amplitude = np.loadtxt(resolve_pulse_shape_file(pulse_shape_file), unpack=True)
    raw = not_a_knot_cubic(knot_x, amplitude)
    fine_t = np.arange(t0, time_end, time_sampling / 100)
    fine_y = evaluate_template(knot_x, raw, fine_t)
    norm = float(np.max(fine_y))
    if not norm > 0:
        raise ValueError('pulse template has no positive sample inside the simulated window')
    fine_y = fine_y / norm
    tau = float(np.sum(0.5 * (fine_y[1:] + fine_y[:-1]) * np.diff(fine_t)))
    coef = not_a_knot_cubic(knot_x, amplitude / norm)

    true_baseline = base + g * rate * (1.0 / (1.0 - xt)) * tau           # :156-157

    n_drop = int(ARTIFICIAL_BACKWARD_TIME // time_sampling)              # :375
    n_sim = len(np.arange(t0, time_end, time_sampling))                  # :197-198
    n_bins = n_sim - n_drop
    if n_bins <= 0:
        raise ValueError('no bins left after dropping the first 40 ns')

    return TraceParams(n_pixels=n_pixels, n_pulses=npe.shape[1], n_samples=n_samples, n_bins=n_bins,
                       n_drop=n_drop, time_start=t0, time_end=time_end, time_sampling=time_sampling,
                       nsb_rate=rate, crosstalk=xt, gain=g, sigma_e=s_e, sigma_1=s_1_rel,
                       baseline=base, n_photon=npe, time_signal=t_sig, jitter=jit,
                       poisson=bool(poisson), sub_binning=int(sub_binning), knot_x=knot_x, coef=coef,
                       norm=norm, tau=tau, true_baseline=true_baseline, set_nsb_rate=rate_set,
                       set_crosstalk=xt_set, set_gain=gain_set)
