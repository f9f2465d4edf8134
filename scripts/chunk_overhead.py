"""What chunking costs the generation kernel: 16384 C4 events generated as launches of n events each,
back to back on one stream, no copies.  Usage: chunk_overhead.py [kernel]   (GPU only)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads

kernel = sys.argv[1] if len(sys.argv) > 1 else 'auto'
E = 16384
g = NTraceGenerator(seed=1, kernel=kernel, **workloads.c4_full_camera())
out = torch.empty((E, 1296, 50), dtype=torch.int16, device='cuda')
for n in (16384, 4096, 1752, 876, 517, 438, 219, 73):
    def run():
        e = 0
        while e < E:
            m = min(n, E - e)
            g.generate(m, first_event=e, out=out[e:e + m])
            e += m
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('%s: %5d events per launch (%3d launches): %.2f ms per %d events' % (g.kernel, n, (E + n - 1) // n, ms, E))
