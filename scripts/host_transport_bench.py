"""dct_generate_host on C4 (or dark) with the plain int16 and the 12-bit transport, pinned destination,
for a range of host thread counts.  Usage: host_transport_bench.py [c4|dark] [threads ...]   (GPU only)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from digicamtoy_b200 import NTraceGenerator, workloads

wl = sys.argv[1] if len(sys.argv) > 1 else 'c4'
threads = [int(x) for x in sys.argv[2:]] or [2, 4, 8, 16]
kw = workloads.c4_full_camera() if wl == 'c4' else workloads.c1_dark(n_photon=3)
E = 16384 if wl == 'c4' else 8192
print('host cores: %d' % os.cpu_count())
ref = None
for transport, thr in [('plain', 0)] + [('packed12', t) for t in threads]:
    os.environ['DCT_HOST_TRANSPORT'] = transport
    if thr:
        os.environ['DCT_HOST_THREADS'] = str(thr)
    g = NTraceGenerator(seed=1, **kw)
    out = torch.empty((E, kw['n_pixels'], g.params.n_bins), dtype=torch.int16).pin_memory()
    for _ in range(2):
        g.generate_host(E, first_event=0, truth=False, out=out)
    if ref is None:
        ref = out.numpy().copy()
    else:
        assert np.array_equal(out.numpy(), ref), 'transports disagree'
    t0 = time.perf_counter()
    for i in range(4):
        g.generate_host(E, first_event=i * E, truth=False, out=out)
    dt = (time.perf_counter() - t0) / 4
    n = E * kw['n_pixels'] * g.params.n_bins
    print('%s %-9s threads %2d: %.2f ms per %d events = %.3e samples/s (%.1f GB/s of int16)' % (
        wl, transport, thr, dt * 1e3, E, n / dt, 2 * n / dt / 1e9))
    g.close()
