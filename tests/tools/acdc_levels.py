"""Throughput of every (NSB, n_photon) level of config_files/ac_dc.yml and of dark.yml on the GPU,
next to the CPU oracle on one core (the table BASELINE.md gives for the reference).  GPU + CPU."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads
from oracle import reference_port as orc

def gpu_events_per_s(kw, E=8192):
    g = NTraceGenerator(seed=1, **kw)
    out = torch.empty((E, g.params.n_pixels, g.params.n_bins), dtype=torch.int16, device='cuda')
    for _ in range(2):
        g.generate(E, first_event=0, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3):
        g.generate(E, first_event=(i + 1) * E, out=out)
    e1.record(); torch.cuda.synchronize()
    return 3 * E / (e0.elapsed_time(e1) * 1e-3), g.params.n_bins

def cpu_s_per_event(kw, n):
    import warnings; warnings.simplefilter('ignore')
    return orc.time_events(n, seed=1, **kw)

rows = []
levels = [('ac_dc', workloads.c2_ac_dc(r, n)) for r in (0, 0.003, 0.04, 0.125, 0.66, 1) for n in (0., 100)]
levels += [('dark', workloads.c1_dark(n)) for n in (0., 3)]
print('| config | NSB GHz | n_photon | GPU events/s | GPU samples/s | CPU oracle s/event (1 core) | CPU samples/s | ratio |')
print('|---|---|---|---|---|---|---|---|')
for name, kw in levels:
    ev, B = gpu_events_per_s(kw)
    heavy = kw['nsb_rate'] >= 0.125
    s = cpu_s_per_event(kw, 1 if heavy else 5)
    print('| %s | %g | %g | %.3e | %.3e | %.3f | %.3e | %.1e |' % (
        name, kw['nsb_rate'], kw['n_photon'], ev, ev * 1296 * B, s, 1296 * B / s, ev * s))
