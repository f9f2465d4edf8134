"""Host logic on CPU: constructor math, spline, C-ABI symbol table."""
import ctypes
import os
import re

import numpy as np
import pytest
from scipy.interpolate import CubicSpline

from oracle import reference_port as orc
from digicamtoy_b200 import native, params

from helpers import GOLDEN_CASES, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_spline_matches_scipy_not_a_knot():
    for f in ('pulse_SST-1M_pixel_0.dat', 'pulse_SST-1M_AfterPreampLowGain.dat'):
        x, y = np.loadtxt(os.path.join(ROOT, 'digicamtoy_b200', 'utils', f), unpack=True)
        c = params.not_a_knot_cubic(x, y)
        ref = CubicSpline(x, y, bc_type='not-a-knot').c      # (4, n-1), highest power first
        scale = np.abs(y).max()
        for k in range(4):
            assert np.allclose(c[:, k], ref[3 - k], rtol=0, atol=1e-9 * scale)
        t = np.linspace(x[0] - 5, x[-1] + 5, 5001)
        want = CubicSpline(x, y)(t)
        want[(t < x[0]) | (t > x[-1])] = 0
        assert np.abs(params.evaluate_template(x, c, t) - want).max() < 1e-12 * scale
        assert params.evaluate_template(x, c, x[-1]) == pytest.approx(y[-1], abs=1e-13)   # inclusive end


@pytest.mark.parametrize('name', [n for n in GOLDEN_CASES if n != 'voltage_drop_refbug'])
def test_constructor_math_matches_reference(name):
    z, meta = load_golden(name)
    p = params.build_params(**meta['kwargs'])
    assert np.allclose(p.true_baseline, z['true_baseline'], rtol=1e-13, atol=0)
    assert p.tau == pytest.approx(float(z['tau']), rel=1e-13)
    assert np.allclose(p.gain, z['gain'], rtol=1e-15)
    assert np.allclose(p.sigma_1, z['sigma_1'], rtol=1e-15)
    assert np.allclose(p.nsb_rate, z['nsb_rate'], rtol=1e-15)
    assert np.allclose(p.crosstalk, z['crosstalk'], rtol=1e-15)
    got = params.evaluate_template(p.knot_x, p.coef, z['template_probe_x'])
    assert np.abs(got - z['template_probe_y']).max() < 1e-13


def test_voltage_drop_is_per_pixel():
    z, meta = load_golden('voltage_drop_refbug')
    p = params.build_params(**meta['kwargs'])
    assert np.allclose(p.gain, z['gain'], rtol=1e-14)           # same polynomials
    assert np.allclose(p.nsb_rate, z['nsb_rate'], rtol=1e-14)
    assert p.n_photon.shape == (5, 1)                           # not the reference's (5, 5)
    assert np.allclose(p.n_photon[:, 0], np.diag(z['n_photon']), rtol=1e-14)


def test_bins_and_ragged_tables():
    p = params.build_params(time_start=0, time_end=100, time_sampling=2, n_pixels=3)
    assert p.n_bins == 50 and p.n_drop == 20 and p.n_samples == 50
    p = params.build_params(n_pixels=3, n_photon=[[100, 133], [60, 60, 60], [100, 133]],
                            time_signal=[[0, 100], [20, 50, 100], [0, 100]], gain=[10, 10, 29])
    assert p.n_photon.shape == (3, 3) and p.n_photon[0, 2] == 0 and p.time_signal[2, 2] == 0
    assert p.jitter.shape == (3, 3)
    p = params.build_params(n_pixels=2, n_photon=np.int64(4))   # numpy scalars accepted
    assert p.n_photon.shape == (2, 1)
    with pytest.raises(ValueError):
        params.build_params(n_pixels=3, gain=[1, 2])
    with pytest.raises(ValueError):
        params.build_params(n_pixels=2, n_photon=[[1, 2], [3, 4]], time_signal=[[1, 2, 3], [1, 2, 3]])
    with pytest.raises(ValueError):
        params.build_params(n_pixels=2, crosstalk=1.2)
    with pytest.raises(FileNotFoundError):
        params.build_params(n_pixels=2, pulse_shape_file='utils/nope.dat')


def test_default_pulse_file_with_leading_slash():
    p = params.build_params(n_pixels=1)                          # reference default '/utils/...'
    assert p.tau == pytest.approx(17.966342626818715, rel=1e-12)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'digicamtoy_b200.h')).read()
    declared = set(re.findall(r'DCT_API\s+[\w\s\*]+?\b(dct_\w+)\s*\(', header))
    assert declared == set(native.EXPORTED_SYMBOLS)
    lib = native.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dct_abi_version() == native.DCT_ABI_VERSION
    assert ctypes.sizeof(native.DctConfig) == 160


def test_shipped_library_is_sm_100a_and_the_tensor_kernel_uses_tcgen05():
    """cuobjdump of the in-tree .so (works without a GPU): the only target is sm_100a, the high-rate
    generation kernel issues tcgen05.mma (SASS UTCHMMA) and commits to mbarriers (UTCBAR), the cube
    leaves through 16-byte streaming stores, and profiles/r02/sass_evidence.txt describes this build."""
    import shutil
    import subprocess
    if shutil.which('cuobjdump') is None:
        pytest.skip('no cuobjdump')
    native.load_library()
    elf = subprocess.run(['cuobjdump', '-lelf', native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_\w+', elf))
    assert archs == {'sm_100a'}, archs
    sass = subprocess.run(['cuobjdump', '-sass', native.LIB_PATH], capture_output=True, text=True).stdout
    kernels = re.split(r'Function : ', sass)[1:]
    by_name = {k.split()[0]: k for k in kernels}
    tensor = [v for k, v in by_name.items() if 'k_generate_tensor' in k]
    assert tensor and all('UTCHMMA' in t and 'UTCBAR' in t and 'STG.E.EF.128' in t for t in tensor)
    sweep = [v for k, v in by_name.items() if 'k_generate_sweep' in k]
    assert sweep and all('FFMA2' in t for t in sweep)
    evidence = open(os.path.join(ROOT, 'profiles', 'r02', 'sass_evidence.txt')).read()
    for name in by_name:                                   # regenerate with scripts/sass_evidence.py after a kernel is added
        short = re.search(r'dct\d+(k_[a-z_]+?)(I|E)', name)
        assert short is None or short.group(1) in evidence, name


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    p = params.build_params(n_pixels=2)
    with pytest.raises(RuntimeError):
        native.Handle(p)
    from digicamtoy_b200 import NTraceGenerator
    with pytest.raises(RuntimeError):
        NTraceGenerator(n_pixels=2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'digicamtoy_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f


def test_analytical_shim_matches_reference_polynomials():
    """reference utils/analytical.py:9-37: cubic polynomials in the NSB rate [Hz], zero from 1.4 GHz."""
    from digicamtoy.utils import analytical as an
    nsb = np.array([0.0, 1e8, 5e8, 1.39e9, 1.4e9, 2e9])
    want_gain = (1.0000 - 4.16866e-10 * nsb + 2.27832e-19 * nsb ** 2 - 6.05033e-29 * nsb ** 3) * (nsb < 1.4e9)
    want_xt = (1. - 7.41234e-10 * nsb + 4.60472e-19 * nsb ** 2 - 1.28336e-28 * nsb ** 3) * (nsb < 1.4e9)
    want_pde = (1. - 2.2929e-10 * nsb + 9.59912e-20 * nsb ** 2 - 2.22249e-29 * nsb ** 3) * (nsb < 1.4e9)
    assert np.allclose(an.true_gain_drop(nsb), want_gain, rtol=1e-14, atol=0)
    assert np.allclose(an.true_xt_drop(nsb), want_xt, rtol=1e-14, atol=0)
    assert np.allclose(an.true_pde_drop(nsb), want_pde, rtol=1e-14, atol=0)
    assert an.true_gain_drop(1.4e9) == 0 and an.true_gain_drop(0.0) == 1.0
    assert np.allclose(an.gain_drop(np.array([0., 1e9])), [1.0, 1 / (1 + 85e-15 * 10e3 * 1e9 * 1e9)])
    z, _ = load_golden('voltage_drop_refbug')      # the reference's own evaluation of the polynomials
    p = params.build_params(n_pixels=5, nsb_rate=0.4, voltage_drop=True, n_photon=20., jitter=1.0)
    assert np.allclose(p.gain, z['gain'], rtol=1e-14) and np.allclose(p.crosstalk, z['crosstalk'], rtol=1e-14)
