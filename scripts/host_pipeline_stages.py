"""dct_generate_host on C4 with one stage left out (a -DDCT_HOST_DEBUG build reads DCT_HOST_SKIP): what the
generation chunks and the copies cost alone and together.  GPU only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads

E = 16384
g = NTraceGenerator(seed=1, **workloads.c4_full_camera())
out = torch.empty((E, 1296, 50), dtype=torch.int16).pin_memory()
for _ in range(2):
    g.generate_host(E, first_event=0, truth=False, out=out)
t0 = time.perf_counter()
for i in range(4):
    g.generate_host(E, first_event=i * E, truth=False, out=out)
dt = (time.perf_counter() - t0) / 4
print('DCT_HOST_SKIP=%s kernel %s: %.2f ms per %d events (%.1f GB/s)' % (os.environ.get('DCT_HOST_SKIP', '-'), g.kernel, dt * 1e3, E, E * 1296 * 100 / dt / 1e9))
if os.environ.get('DCT_HOST_SKIP', 'none') == 'none':
    # the plain copies on the same box: one 2.1 GB cudaMemcpyAsync, and the same bytes in 113 MB pieces
    dev = torch.empty((E, 1296, 50), dtype=torch.int16, device='cuda')
    for pieces in (1, 19, 75):
        n = (E + pieces - 1) // pieces
        def run():
            for e in range(0, E, n):
                out[e:e + n].copy_(dev[e:e + n], non_blocking=True)
        run(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            run()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        print('plain D2H in %d piece(s): %.2f ms (%.1f GB/s)' % (pieces, dt * 1e3, E * 1296 * 100 / dt / 1e9))
