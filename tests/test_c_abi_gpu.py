"""The boundary is a real C ABI: a plain-C program (gcc, no CUDA headers, no Python) links the
library, generates a cube into a malloc'ed buffer and must get the bits the Python path gets."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_consumer_matches_python_path(tmp_path):
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc on this box')
    from digicamtoy_b200 import NTraceGenerator, native
    exe = str(tmp_path / 'c_abi_demo')
    libdir = os.path.join(ROOT, 'digicamtoy_b200')
    subprocess.run([gcc, '-O2', '-o', exe, os.path.join(ROOT, 'examples', 'c_abi_demo.c'),
                    '-I', os.path.join(ROOT, 'include'), '-L', libdir, '-l:libdigicamtoy_sm100.so',
                    '-Wl,-rpath,' + libdir], check=True)
    kw = dict(n_pixels=77, nsb_rate=0.4, n_photon=12., sigma_e=1.25, gain=5, baseline=280, jitter=0.2)
    g = NTraceGenerator(seed=2468, **kw)
    cfg = str(tmp_path / 'config.bin')
    native.export_config(g.params, cfg)
    out = subprocess.run([exe, cfg, '2468', '33'], check=True, capture_output=True, text=True).stdout
    fields = dict(kv.split('=') for kv in out.split())
    adc = g.generate(33).cpu().numpy().view(np.uint16).ravel()
    assert int(fields['samples']) == adc.size
    assert int(fields['sum']) == int(adc.sum(dtype=np.uint64))
    mix = 1469598103934665603
    for v in adc.tolist():
        mix = ((mix ^ v) * 1099511628211) & (2 ** 64 - 1)
    assert int(fields['fnv']) == mix
    assert int(fields['abi']) == native.DCT_ABI_VERSION
