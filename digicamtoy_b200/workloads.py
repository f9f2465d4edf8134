"""Named workloads of BASELINE.json / SURVEY.md section 8(d), as NTraceGenerator keyword sets."""
from __future__ import annotations

import os

import numpy as np

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C4_SEED = 20261019


def load_yaml(path):
    import yaml
    loader = getattr(yaml, 'CSafeLoader', yaml.SafeLoader)
    with open(path) as stream:
        return yaml.load(stream, Loader=loader)


def c4_full_camera(n_pixels=1296):
    """C4: full camera, 50 bins, 1 GHz NSB on every pixel, one signal pulse per pixel whose size
    rises from 1 to 3162 pe across the camera so that the top ~10 % of pixels saturate 12 bits."""
    p = np.arange(n_pixels)
    return dict(time_start=0, time_end=200, time_sampling=4, n_pixels=n_pixels, nsb_rate=1.0,
                crosstalk=0.08, gain_nsb=True,
                n_photon=[[float(10 ** (3.5 * i / max(n_pixels - 1, 1)))] for i in p],
                poisson=True, sigma_e=1.25, sigma_1=0.8, gain=5, baseline=280, time_signal=20,
                jitter=0.2, pulse_shape_file='utils/pulse_SST-1M_pixel_0.dat')


def c1_dark(n_photon=0.):
    cfg = load_yaml(os.path.join(REPO_ROOT, 'config_files', 'dark.yml'))
    return _generator_kwargs(cfg, cfg['nsb_rate'][0], n_photon)


def c2_ac_dc(nsb_rate, n_photon):
    cfg = load_yaml(os.path.join(REPO_ROOT, 'config_files', 'ac_dc.yml'))
    return _generator_kwargs(cfg, nsb_rate, n_photon)


def _generator_kwargs(cfg, nsb_rate, n_photon):
    kw = {k: v for k, v in cfg.items() if k not in ('output_directory', 'file_basename', 'events_per_level')}
    kw['nsb_rate'] = nsb_rate
    kw['n_photon'] = n_photon
    return kw


def shard_events(n_events, rank, world_size):
    """Contiguous range of the global event index owned by `rank`: the
    Philox counter carries the global index, so any sharding reproduces the single-GPU cube."""
    base, extra = divmod(int(n_events), int(world_size))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)
