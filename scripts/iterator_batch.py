"""Iterator rate (for event in NTraceGenerator(...)) against the size of its pinned batches.  GPU only.
Usage: iterator_batch.py [c4|dark]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from digicamtoy_b200 import NTraceGenerator, workloads

wl = sys.argv[1] if len(sys.argv) > 1 else 'c4'
kw = workloads.c4_full_camera() if wl == 'c4' else workloads.c1_dark(n_photon=3)
for be in (None, 258, 517, 1035, 2070, 4140):
    N = 3 * (be or 2070) + 4000
    W = 2 * (be or 300) + 50                      # both pinned slots exist before the clock starts
    g = NTraceGenerator(seed=1, n_events=N + W, batch_events=be, **kw)
    it = iter(g)
    for _ in range(W):
        next(it)
    s = 0
    t0 = time.perf_counter()
    for _ in range(N - 1):
        ev = next(it)
        s += int(ev.adc_count[0, 0])
    dt = time.perf_counter() - t0
    print('%s batch_events %s (%d): %.3e events/s  (%.2f us per event)' % (wl, be, g.batch_events, (N - 1) / dt, dt / (N - 1) * 1e6))
    g.close()
