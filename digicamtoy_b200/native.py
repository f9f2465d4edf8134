"""ctypes binding of ``libdigicamtoy_sm100.so`` (C ABI in ``include/digicamtoy_b200.h``).

There is no CPU implementation behind this module: if the library is missing and cannot be
built, or no CUDA device is present, calls raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

import numpy as np

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = 'libdigicamtoy_sm100.so'
LIB_PATH = os.environ.get('DCT_LIB_PATH') or os.path.join(PACKAGE_DIR, LIB_NAME)    # (override: timing experiments against older builds)
CSRC_DIR = os.path.join(PACKAGE_DIR, 'csrc')

DCT_ABI_VERSION = 3
KERNEL_AUTO, KERNEL_SCATTER, KERNEL_SWEEP, KERNEL_GRID, KERNEL_TENSOR = 0, 1, 2, 3, 4
KERNEL_NAMES = {KERNEL_AUTO: 'auto', KERNEL_SCATTER: 'scatter', KERNEL_SWEEP: 'sweep', KERNEL_GRID: 'grid',
                KERNEL_TENSOR: 'tensor'}

EXPORTED_SYMBOLS = (
    'dct_abi_version', 'dct_last_error', 'dct_create', 'dct_destroy', 'dct_selected_kernel',
    'dct_generate', 'dct_generate_stats', 'dct_generate_analog', 'dct_generate_host', 'dct_replay_f64', 'dct_replay_f32',
    'dct_stats_accumulate', 'dct_write_calibrate', 'dct_test_sample', 'dct_trace_pe', 'dct_test_unpack12', 'dct_test_chunk_schedule',
)

_dp = ctypes.POINTER(ctypes.c_double)


class DctConfig(ctypes.Structure):
    """Mirror of ``struct dct_config``."""
    _fields_ = [
        ('struct_size', ctypes.c_uint32), ('n_pixels', ctypes.c_uint32),
        ('n_pulses', ctypes.c_uint32), ('n_bins', ctypes.c_uint32), ('n_drop', ctypes.c_uint32),
        ('n_knots', ctypes.c_uint32), ('poisson', ctypes.c_uint32), ('sub_binning', ctypes.c_uint32),
        ('adc_min', ctypes.c_int32), ('adc_max', ctypes.c_int32), ('kernel', ctypes.c_uint32),
        ('reserved', ctypes.c_uint32),
        ('t0', ctypes.c_double), ('t_end', ctypes.c_double), ('dt', ctypes.c_double),
        ('nsb_rate', _dp), ('crosstalk', _dp), ('gain', _dp), ('sigma_1', _dp), ('sigma_e', _dp),
        ('baseline', _dp), ('n_photon', _dp), ('time_signal', _dp), ('jitter', _dp),
        ('knot_x', _dp), ('coef', _dp),
    ]


def build_library(verbose=False):
    """Compile the CUDA sources for sm_100a into the package directory (in-tree)."""
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; cannot build ' + LIB_NAME)
    cmd = ['make', '-C', CSRC_DIR, 'NVCC=' + nvcc]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('building {} failed:\n{}'.format(LIB_NAME, res.stderr[-4000:]))
    return LIB_PATH


_lib = None


def load_library():
    """Load (building first if absent) and type the C ABI.  Never returns a fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build_library()
    lib = ctypes.CDLL(LIB_PATH)
    vp, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int32
    lib.dct_abi_version.restype = ctypes.c_int
    lib.dct_last_error.restype = ctypes.c_char_p
    lib.dct_create.argtypes = [ctypes.POINTER(DctConfig), ctypes.c_int, ctypes.POINTER(vp)]
    lib.dct_destroy.argtypes = [vp]
    lib.dct_destroy.restype = None
    lib.dct_selected_kernel.argtypes = [vp]
    lib.dct_generate.argtypes = [vp, u64, u64, u32, vp, vp, vp, vp]
    lib.dct_generate_host.argtypes = [vp, u64, u64, u32, vp, vp, vp]
    if hasattr(lib, 'dct_generate_analog'):
        lib.dct_generate_analog.argtypes = [vp, u64, u64, u32, vp, vp, vp]
    for f in (lib.dct_replay_f64, lib.dct_replay_f32):
        f.argtypes = [vp, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    if hasattr(lib, 'dct_generate_stats'):          # (absent only in older builds loaded through DCT_LIB_PATH)
        lib.dct_generate_stats.argtypes = [vp, u64, u64, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, u32, vp, vp]
    lib.dct_stats_accumulate.argtypes = [vp, vp, u32, vp, vp, vp, vp, vp, vp, i32, u32, vp, vp]
    lib.dct_write_calibrate.argtypes = [vp, ctypes.c_size_t, vp]
    lib.dct_trace_pe.argtypes = [vp, u64, u64, u32, u32, vp, vp, vp, vp, vp]
    lib.dct_test_sample.argtypes = [u32, ctypes.c_double, ctypes.c_double, u64, u32, vp, vp]
    if hasattr(lib, 'dct_test_unpack12'):
        lib.dct_test_unpack12.argtypes = [vp, vp, u64, u32]
    if hasattr(lib, 'dct_test_chunk_schedule'):
        lib.dct_test_chunk_schedule.argtypes = [u32, u32, vp, u32]
    for name in EXPORTED_SYMBOLS:
        if name not in ('dct_last_error', 'dct_destroy') and hasattr(lib, name):
            getattr(lib, name).restype = ctypes.c_int
    if lib.dct_abi_version() != DCT_ABI_VERSION and not os.environ.get('DCT_LIB_PATH'):
        raise RuntimeError('{} has ABI version {}, expected {}'.format(
            LIB_PATH, lib.dct_abi_version(), DCT_ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load_library().dct_last_error().decode('utf-8', 'replace')
        if rc in (-1, -3):
            raise ValueError('digicamtoy_b200: ' + msg)
        raise RuntimeError('digicamtoy_b200 (code {}): {}'.format(rc, msg))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Handle:
    """Owns one ``dct_handle`` (device parameter block for one generator configuration)."""

    def __init__(self, params, device=0, adc_min=0, adc_max=4095, kernel=KERNEL_AUTO):
        self.lib = load_library()
        self._keep = dict(
            nsb_rate=_f64(params.nsb_rate), crosstalk=_f64(params.crosstalk), gain=_f64(params.gain),
            sigma_1=_f64(params.sigma_1), sigma_e=_f64(params.sigma_e), baseline=_f64(params.baseline),
            n_photon=_f64(params.n_photon), time_signal=_f64(params.time_signal),
            jitter=_f64(params.jitter), knot_x=_f64(params.knot_x), coef=_f64(params.coef))
        cfg = DctConfig()
        cfg.struct_size = ctypes.sizeof(DctConfig)
        cfg.n_pixels, cfg.n_pulses = params.n_pixels, params.n_pulses
        cfg.n_bins, cfg.n_drop, cfg.n_knots = params.n_bins, params.n_drop, len(params.knot_x)
        cfg.poisson, cfg.sub_binning = int(params.poisson), max(0, int(params.sub_binning))
        cfg.adc_min, cfg.adc_max, cfg.kernel = int(adc_min), int(adc_max), int(kernel)
        cfg.t0, cfg.t_end, cfg.dt = float(params.time_start), float(params.time_end), float(
            params.time_sampling)
        for name, arr in self._keep.items():
            setattr(cfg, name, arr.ctypes.data_as(_dp))
        self.ptr = ctypes.c_void_p()
        self.device = int(device)
        self.adc_min, self.adc_max = int(adc_min), int(adc_max)
        check(self.lib.dct_create(ctypes.byref(cfg), self.device, ctypes.byref(self.ptr)))
        self.kernel = KERNEL_NAMES.get(self.lib.dct_selected_kernel(self.ptr), '?')

    def close(self):
        # forget the pointer BEFORE destroying it: during interpreter shutdown anything after the C
        # call may raise (module globals are already gone), and a second close() -- the generator's
        # __del__ and then this object's -- must not destroy the handle again
        ptr, self.ptr = getattr(self, 'ptr', None), None
        if ptr is not None and ptr.value:
            self.lib.dct_destroy(ptr)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- thin wrappers; every pointer is an integer address (torch ``data_ptr()`` or numpy) -----
    def generate(self, seed, first_event, n_events, adc_ptr, sig_time_ptr=None, sig_charge_ptr=None,
                 stream=None):
        check(self.lib.dct_generate(self.ptr, seed, first_event, n_events, adc_ptr, sig_time_ptr,
                                    sig_charge_ptr, stream))

    def generate_stats(self, seed, first_event, n_events, adc_ptr, sig_time_ptr=None, sig_charge_ptr=None,
                       adc_hist=None, bin_sum=None, bin_sumsq=None, pixel_sum=None, pixel_sumsq=None,
                       charge_hist=None, charge_lo=0, n_charge_bins=0, moments=None, stream=None):
        check(self.lib.dct_generate_stats(self.ptr, seed, first_event, n_events, adc_ptr, sig_time_ptr,
                                          sig_charge_ptr, adc_hist, bin_sum, bin_sumsq, pixel_sum, pixel_sumsq,
                                          charge_hist, charge_lo, n_charge_bins, moments, stream))

    def generate_analog(self, seed, first_event, n_events, adc_ptr, analog_ptr, stream=None):
        check(self.lib.dct_generate_analog(self.ptr, seed, first_event, n_events, adc_ptr, analog_ptr, stream))

    def generate_host(self, seed, first_event, n_events, adc_ptr, sig_time_ptr=None,
                      sig_charge_ptr=None):
        check(self.lib.dct_generate_host(self.ptr, seed, first_event, n_events, adc_ptr,
                                         sig_time_ptr, sig_charge_ptr))

    def replay(self, precision, n_events, offsets_ptr, time_ptr, amp_ptr, sig_time_ptr, sig_amp_ptr,
               noise_ptr, adc_ptr, clamped_ptr=None, stream=None):
        fn = self.lib.dct_replay_f64 if precision == 'f64' else self.lib.dct_replay_f32
        check(fn(self.ptr, n_events, offsets_ptr, time_ptr, amp_ptr, sig_time_ptr, sig_amp_ptr,
                 noise_ptr, adc_ptr, clamped_ptr, stream))

    def trace_pe(self, seed, first_event, n_events, max_pe, n_pe_ptr, pe_time_ptr=None, pe_amp_ptr=None,
                 noise_ptr=None, stream=None):
        check(self.lib.dct_trace_pe(self.ptr, seed, first_event, n_events, max_pe, n_pe_ptr, pe_time_ptr,
                                    pe_amp_ptr, noise_ptr, stream))

    def stats_accumulate(self, adc_ptr, n_events, adc_hist=None, bin_sum=None, bin_sumsq=None,
                         pixel_sum=None, pixel_sumsq=None, charge_hist=None, charge_lo=0,
                         n_charge_bins=0, moments=None, stream=None):
        check(self.lib.dct_stats_accumulate(self.ptr, adc_ptr, n_events, adc_hist, bin_sum, bin_sumsq,
                                            pixel_sum, pixel_sumsq, charge_hist, charge_lo,
                                            n_charge_bins, moments, stream))


def write_calibrate(buf_ptr, n, stream=None):
    check(load_library().dct_write_calibrate(buf_ptr, n, stream))


SAMPLERS = {'normal': 0, 'exp_gap': 1, 'borel': 2, 'poisson': 3, 'branching': 4, 'uniform': 5,
            'poisson_top': 6, 'borel_top': 7, 'poisson_inv_top': 8, 'poisson_table_top': 9,
            'normal_radius_top': 10}


def test_sample(kind, a, b, seed, n, out_ptr, stream=None):
    check(load_library().dct_test_sample(SAMPLERS[kind], float(a), float(b), seed, n, out_ptr, stream))


def test_unpack12(packed, n_samples, threads=0, offset=0):
    """Host-only test hook: NumPy uint8 stream of 12-bit samples -> int16 array of n_samples
    (``offset``: misalign the destination by that many int16)."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    out = np.empty(int(n_samples) + 32, dtype=np.int16)[int(offset):int(offset) + int(n_samples)]
    check(load_library().dct_test_unpack12(packed.ctypes.data, out.ctypes.data, int(n_samples), int(threads)))
    return out


def test_chunk_schedule(unit, n_events):
    """Host-only test hook: chunk ends [events] of a dct_generate_host call."""
    ends = np.zeros(1 << 16, dtype=np.uint32)
    n = load_library().dct_test_chunk_schedule(int(unit), int(n_events), ends.ctypes.data, len(ends))
    if n < 0:
        raise RuntimeError(load_library().dct_last_error().decode())
    return ends[:n].copy()


def export_config(params, path, adc_min=0, adc_max=4095, kernel=KERNEL_AUTO):
    """Flat binary dump of a dct_config for non-Python consumers (examples/c_abi_demo.c):
    10 x uint32 scalars, 3 x float64 (t0, t_end, dt), then every array as float64."""
    head = np.array([params.n_pixels, params.n_pulses, params.n_bins, params.n_drop, len(params.knot_x),
                     int(params.poisson), max(0, int(params.sub_binning)), np.int32(adc_min).astype(np.uint32),
                     np.int32(adc_max).astype(np.uint32), kernel], dtype=np.uint32)
    with open(path, 'wb') as f:
        f.write(head.tobytes())
        f.write(np.array([params.time_start, params.time_end, params.time_sampling], dtype=np.float64).tobytes())
        for a in (params.nsb_rate, params.crosstalk, params.gain, params.sigma_1, params.sigma_e,
                  params.baseline, params.n_photon, params.time_signal, params.jitter, params.knot_x,
                  params.coef):
            f.write(_f64(a).tobytes())
