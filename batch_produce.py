#!/usr/bin/env python
"""One launcher for the reference's commissioning production.

The reference produces its commissioning set in three steps: ``config_files/generate.py`` writes
one YAML per (n_photon level i, NSB level j, file id k) -- ``ac_{i}_dc_{j}_id_{k}.yml``, 500 x 1 x 10
of them by default (generate.py:41-71) -- ``jobs.sh`` submits one SLURM job per file
(jobs.sh:3-16), and each job runs ``produce_data.py -y <that yaml>`` on one core for up to two
hours (run.sh:3-20).

Here the same (i, j, k) grid is walked by one process per GPU:

    python batch_produce.py --output_directory /data/mc                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 batch_produce.py ...  # jobs dealt round-robin

Each job is one call of ``produce_data.produce_level`` with exactly the option dictionary that
generate.py would have dumped (same keys, same values, same file names), so the files are what the
5000 SLURM jobs would have written.  Compression and file writing run on a small thread pool
(``--writers``) while the GPU generates the next file; ``--max_in_flight`` bounds host memory.

``--write_configs DIR`` only emits the YAML files (the generate.py step), ``--dry_run`` lists the
jobs of this rank.  Ranges are inclusive, like AC_START..AC_END in jobs.sh.
"""
import argparse
import logging
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

OUTPUT_FILE_NAME = 'ac_{}_dc_{}_id_{}.hdf5'          # generate.py:50
CONFIG_FILE_NAME = 'ac_{}_dc_{}_id_{}.yml'           # generate.py:51 (below commissioning/)


def standard_parameters(n_pixels=1296, events_per_level=1000, output_directory='/sst1m/MC/digicamtoy/'):
    """The option dictionary of generate.py:8-47 without the light levels."""
    p = dict()
    p['output_directory'] = output_directory
    p['events_per_level'] = int(events_per_level)
    # camera (generate.py:21-30)
    p['time_start'] = 0
    p['time_end'] = 368
    p['time_sampling'] = 4
    p['n_pixels'] = int(n_pixels)
    p['gain_nsb'] = True
    p['poisson'] = True
    p['pulse_shape_file'] = 'utils/pulse_SST-1M_pixel_0.dat'
    # SiPM, one value per pixel (generate.py:8-18)
    for key, val in (('sigma_e', 0.8), ('sigma_1', 0.8), ('gain', 5.8), ('baseline', 200),
                     ('time_signal', 20), ('jitter', 1.155), ('crosstalk', 0.08)):
        p[key] = [val] * p['n_pixels']
    return p


def light_levels(n_photon_spec=(0, 500, 1), nsb_spec=(0.003, 2.0, 30)):
    """(n_photons, nsb_rates) of generate.py:53-54: arange(start, stop, step), linspace(lo, hi, num)."""
    a0, a1, da = n_photon_spec
    lo, hi, num = nsb_spec
    return np.arange(a0, a1, da), np.linspace(lo, hi, num=int(num))


def job_parameters(base, n_photons, nsb_rates, i, j, k):
    """Options of job (i, j, k): generate.py:33-38,64-66."""
    p = dict(base)
    p['nsb_rate'] = [float(nsb_rates[j])]
    p['n_photon'] = [float(n_photons[i])]
    p['file_basename'] = OUTPUT_FILE_NAME.format(i, j, k)
    return p


def job_grid(ac, dc, file_id):
    """(i, j, k) in the submission order of jobs.sh:10-16; all three ranges inclusive."""
    return [(i, j, k) for i in range(ac[0], ac[1] + 1) for j in range(dc[0], dc[1] + 1)
            for k in range(file_id[0], file_id[1] + 1)]


def deal(jobs, rank, world):
    return jobs[rank::world]


def job_seed(base_seed, n_dc, n_ids, i, j, k):
    """Distinct per file and independent of which rank runs it."""
    return None if base_seed is None else int(base_seed) + (i * n_dc + j) * n_ids + k


def write_configs(directory, base, n_photons, nsb_rates, jobs):
    import yaml
    dumper = getattr(yaml, 'CDumper', yaml.Dumper)
    os.makedirs(directory, exist_ok=True)
    paths = []
    for i, j, k in jobs:
        path = os.path.join(directory, CONFIG_FILE_NAME.format(i, j, k))
        with open(path, mode='w') as f:
            yaml.dump(job_parameters(base, n_photons, nsb_rates, i, j, k), f, Dumper=dumper)
        paths.append(path)
    return paths


def _range(text):
    a, b = text.split(':')
    return int(a), int(b)


def _spec(text, kinds):
    parts = text.split(':')
    if len(parts) != 3:
        raise argparse.ArgumentTypeError('expected three values separated by ":"')
    return tuple(kind(x) for kind, x in zip(kinds, parts))


def parse(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--ac', type=_range, default=(0, 499), help='n_photon level indices START:END (jobs.sh AC_*)')
    ap.add_argument('--dc', type=_range, default=(0, 0), help='NSB level indices START:END (jobs.sh DC_*)')
    ap.add_argument('--file_id', type=_range, default=(0, 9), help='file ids START:END (jobs.sh FILE_ID_*)')
    ap.add_argument('--n_photon', type=lambda t: _spec(t, (float, float, float)), default=(0, 500, 1),
                    help='START:STOP:STEP of the n_photon levels (generate.py:54)')
    ap.add_argument('--nsb_rate', type=lambda t: _spec(t, (float, float, int)), default=(0.003, 2.0, 30),
                    help='LOW:HIGH:NUM of the NSB levels in GHz (generate.py:53)')
    ap.add_argument('--events_per_level', type=int, default=1000)
    ap.add_argument('--n_pixels', type=int, default=1296)
    ap.add_argument('--output_directory', default='/sst1m/MC/digicamtoy/')
    ap.add_argument('--seed', type=int, default=None, help='base seed; file (i,j,k) gets a distinct offset')
    ap.add_argument('--device', type=int, default=None)
    ap.add_argument('--writers', type=int, default=max(1, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else 4),
                    help='threads that compress and write files (default: one per host core of this rank)')
    ap.add_argument('--max_in_flight', type=int, default=None,
                    help='files generated but not yet written (default: writers + 2)')
    ap.add_argument('--resume', action='store_true', help='skip files that already exist instead of refusing to run')
    ap.add_argument('--compress_level', type=int, default=None,
                    help='zlib level of the .npz bundles (default 4, what h5py uses for compression="gzip")')
    ap.add_argument('--write_configs', metavar='DIR', default=None, help='only write the YAML files there')
    ap.add_argument('--dry_run', action='store_true')
    ap.add_argument('-d', '--debug', action='store_true')
    return ap.parse_args(argv)


def main(argv=None):
    a = parse(argv)
    log = logging
    log.getLogger().setLevel(logging.DEBUG if a.debug else logging.INFO)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    device = a.device if a.device is not None else int(os.environ.get('LOCAL_RANK', '0'))

    n_photons, nsb_rates = light_levels(a.n_photon, a.nsb_rate)
    if not (0 <= a.ac[0] <= a.ac[1] < len(n_photons)) or not (0 <= a.dc[0] <= a.dc[1] < len(nsb_rates)) \
            or not (0 <= a.file_id[0] <= a.file_id[1]):
        raise ValueError('level range outside the %d n_photon x %d NSB levels' % (len(n_photons), len(nsb_rates)))
    base = standard_parameters(a.n_pixels, a.events_per_level, a.output_directory)
    jobs = job_grid(a.ac, a.dc, a.file_id)
    mine = deal(jobs, rank, world)

    if a.write_configs is not None:
        paths = write_configs(a.write_configs, base, n_photons, nsb_rates, mine)
        log.info('rank %d: wrote %d configuration files to %s', rank, len(paths), a.write_configs)
        return paths
    if a.dry_run:
        for i, j, k in mine:
            print(os.path.join(a.output_directory, OUTPUT_FILE_NAME.format(i, j, k)))
        return [os.path.join(a.output_directory, OUTPUT_FILE_NAME.format(i, j, k)) for i, j, k in mine]

    import produce_data
    # the reference refuses to overwrite (produce_data.py:47-57): check every file of this rank first.
    # --resume is the reference's way of finishing a production -- re-submitting the missing file ids
    # (run.sh:19) -- in one flag: files that exist are skipped (a file only gets its name once complete).
    todo = []
    for i, j, k in mine:
        f = os.path.join(a.output_directory, OUTPUT_FILE_NAME.format(i, j, k))
        if produce_data.output_exists(f):
            if not a.resume:
                log.error('File {} already exists'.format(f))
                return 1
            continue
        todo.append((i, j, k))
    if a.resume:
        log.info('rank %d: %d of %d files already there, %d to do', rank, len(mine) - len(todo), len(mine), len(todo))
    mine = todo
    os.makedirs(a.output_directory, exist_ok=True)
    if a.compress_level is not None:
        produce_data.COMPRESS_LEVEL = a.compress_level
    n_dc, n_ids = len(nsb_rates), a.file_id[1] + 1
    slots = threading.Semaphore(max(1, a.max_in_flight if a.max_in_flight is not None else a.writers + 2))
    pool = ThreadPoolExecutor(max_workers=max(1, a.writers))
    futures = []

    def finish(close):
        def run():
            try:
                return close()
            finally:
                slots.release()
        futures.append(pool.submit(run))

    t0 = time.perf_counter()
    t_gen = 0.0
    for i, j, k in mine:
        slots.acquire()
        config = job_parameters(base, n_photons, nsb_rates, i, j, k)
        # produce_data walks the one-element level lists and hands the generator scalars (:71-93)
        config['nsb_rate'], config['n_photon'] = config['nsb_rate'][0], config['n_photon'][0]
        filename = os.path.join(a.output_directory, config['file_basename'].format(0))   # produce_data.py:63
        t1 = time.perf_counter()
        try:
            produce_data.produce_level(config, filename, a.events_per_level,
                                       job_seed(a.seed, n_dc, n_ids, i, j, k), device, finish=finish)
        except BaseException:
            slots.release()
            pool.shutdown(wait=True)
            raise
        t_gen += time.perf_counter() - t1
    written = [f.result() for f in futures]
    pool.shutdown(wait=True)
    wall = time.perf_counter() - t0
    n_events = len(mine) * a.events_per_level
    log.info('rank %d/%d: %d files, %d events in %.1f s (%.0f events/s; generation + D2H %.1f s, the rest is '
             'compression and file writing)', rank, world, len(mine), n_events, wall,
             n_events / wall if wall > 0 else 0.0, t_gen)
    return written


if __name__ == '__main__':
    out = main()
    sys.exit(out if isinstance(out, int) else 0)
