"""Error behaviour of the boundary: C side returns codes + thread-local message, Python raises
ValueError for bad configuration and RuntimeError for CUDA problems; nothing aborts."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c_abi_rejects_bad_config_without_aborting():
    from digicamtoy_b200 import native
    from digicamtoy_b200.params import build_params
    lib = native.load_library()
    h = ctypes.c_void_p()
    assert lib.dct_create(None, 0, ctypes.byref(h)) == -1 and b'NULL' in lib.dct_last_error()
    cfg = native.DctConfig()
    cfg.struct_size = 8                                       # ABI guard
    assert lib.dct_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == -1
    assert b'struct_size' in lib.dct_last_error() and not h.value
    p = build_params(n_pixels=4)
    with pytest.raises(ValueError):
        native.Handle(p, device=0, adc_min=10, adc_max=5)     # inverted saturation range
    with pytest.raises(ValueError):
        native.Handle(p, device=99)                           # no such device
    good = native.Handle(p, device=0)
    with pytest.raises(ValueError):
        good.generate(1, 0, 4, None)                          # NULL output
    assert lib.dct_generate(good.ptr, 1, 0, 0, None, None, None, None) in (0, -1)   # n_events = 0: no launch
    good.close()
    good.close()                                              # idempotent


def test_python_surface_raises_value_errors():
    from digicamtoy_b200 import NTraceGenerator
    with pytest.raises(ValueError):
        NTraceGenerator(n_pixels=4, crosstalk=1.5)
    with pytest.raises(ValueError):
        NTraceGenerator(n_pixels=4, gain=[1, 2, 3])
    with pytest.raises(ValueError):
        NTraceGenerator(n_pixels=4, kernel='nope')
    with pytest.raises(ValueError):                           # sweep kernel only exists for the 22-tap / 8-phase table
        NTraceGenerator(n_pixels=4, kernel='sweep', pulse_shape_file='utils/pulse_SST-1M_AfterPreampLowGain.dat')
    with pytest.raises(ValueError):                           # grid kernel only exists for sub_binning > 0
        NTraceGenerator(n_pixels=4, kernel='grid')
    with pytest.raises(ValueError):
        NTraceGenerator(n_pixels=4, device='cpu')
    with pytest.raises(FileNotFoundError):
        NTraceGenerator(n_pixels=4, pulse_shape_file='utils/missing.dat')
    g = NTraceGenerator(n_pixels=4, seed=1)
    import torch
    with pytest.raises(ValueError):
        g.generate(2, out=torch.empty(7, dtype=torch.int16, device='cuda'))
    with pytest.raises(ValueError):
        g.generate(2, out=torch.empty((2, 4, 50), dtype=torch.float32, device='cuda'))


def test_many_handles_and_streams():
    """Handles are independent and launches respect the caller's stream."""
    import torch
    from digicamtoy_b200 import NTraceGenerator
    gens = [NTraceGenerator(n_pixels=64, nsb_rate=0.2 * (i + 1), seed=i) for i in range(4)]
    streams = [torch.cuda.Stream() for _ in gens]
    outs = []
    for g, s in zip(gens, streams):
        with torch.cuda.stream(s):
            outs.append(g.generate(50))
    torch.cuda.synchronize()
    for i, (g, o) in enumerate(zip(gens, outs)):
        again = g.generate(50, first_event=0)
        torch.cuda.synchronize()
        assert torch.equal(o, again)
    assert outs[0].float().mean() < outs[3].float().mean()     # more NSB, higher pedestal


def test_process_exits_cleanly_without_an_explicit_close():
    """A generator that is never closed is destroyed once at interpreter shutdown -- by its own
    __del__ or by its handle's, whichever runs first, never by both (round 2 found a double
    dct_destroy there: glibc reported 'double free or corruption' at exit)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for kernel in ('sweep', 'tensor'):
        out = subprocess.run([sys.executable, os.path.join(root, 'scripts', 'exit_check.py'), kernel, 'noclose', '2'],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-1000:]
        assert 'done ' + kernel in out.stdout
        assert 'double free' not in out.stderr and 'corruption' not in out.stderr, out.stderr[-1000:]


def test_calls_after_close_raise_instead_of_touching_freed_memory():
    import torch
    from digicamtoy_b200 import NTraceGenerator
    g = NTraceGenerator(n_pixels=8, nsb_rate=0.1, seed=1)
    g.generate(2)
    torch.cuda.synchronize()
    g.close()
    g.close()                                   # idempotent
    with pytest.raises(ValueError):
        g.generate(2)
