#!/usr/bin/env python
"""Drop-in for the reference's ``produce_data.py`` (reference produce_data.py:16-147).

Same command line (``-y/--yaml_config``, ``-d/--debug``), same YAML schema (generator keyword
arguments + ``output_directory``, ``file_basename``, ``events_per_level``), same refusal to
overwrite, same NSB-major file numbering, same datasets:

    config/<every option>, config/date
    data/true_baseline (P,), data/adc_count uint16 (E,P,B), data/time (E,P,K), data/charge (E,P,K)

but every level is generated on the GPU in one chunked call (``NTraceGenerator.generate_host``,
D2H overlapped with generation) instead of an event-by-event Python loop.  Output is HDF5 when
h5py is importable; otherwise the identical logical layout goes into ``<file>.npz`` (keys are the
dataset paths).  Launched under torchrun, the (nsb, n_photon) levels are dealt round-robin to the
ranks -- the replacement for the reference's SLURM job array (jobs.sh, run.sh).

Additions (all optional): ``--seed``, ``--output_directory``, ``--events_per_level``, ``--device``.
"""
import copy
import datetime
import logging
import os
import sys
from optparse import OptionParser

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def natural_size(n):
    for unit in ('Bytes', 'KiB', 'MiB', 'GiB', 'TiB'):
        if n < 1024 or unit == 'TiB':
            return ('%d %s' % (n, unit)) if unit == 'Bytes' else ('%.1f %s' % (n, unit))
        n /= 1024.0


def load_options(argv=None):
    parser = OptionParser()
    parser.add_option('-y', '--yaml_config', dest='yaml_config',
                      help='full path of the yaml configuration function', default='options/ac_dc_scan.yaml')
    parser.add_option('-d', '--debug', action='store_true', dest='debug', default=False, help='move to debug')
    parser.add_option('--seed', dest='cli_seed', type='int', default=None, help='base seed (level i uses seed+i)')
    parser.add_option('--output_directory', dest='cli_output_directory', default=None)
    parser.add_option('--events_per_level', dest='cli_events_per_level', type='int', default=None)
    parser.add_option('--device', dest='cli_device', type='int', default=None)
    options, _ = parser.parse_args(argv)
    import yaml
    loader = getattr(yaml, 'CSafeLoader', yaml.SafeLoader)
    with open(options.yaml_config) as stream:
        for key, val in yaml.load(stream, Loader=loader).items():
            options.__dict__[key] = val                                   # produce_data.py:39-45
    cli = {k: options.__dict__.pop(k) for k in list(options.__dict__) if k.startswith('cli_')}
    if cli['cli_output_directory'] is not None:
        options.output_directory = cli['cli_output_directory']
    if cli['cli_events_per_level'] is not None:
        options.events_per_level = cli['cli_events_per_level']
    return options, cli


def level_filenames(options):
    n = len(options.n_photon) * len(options.nsb_rate)
    return [os.path.join(options.output_directory, options.file_basename.format(i)) for i in range(n)]


COMPRESS_LEVEL = 4        # zlib level of the .npz bundle = h5py's default for compression='gzip' (reference :129-142)


class Writer:
    """HDF5 through h5py when present, else an .npz bundle with the dataset paths as keys.  Either
    way the file is written under a temporary name and renamed when complete, so an interrupted
    run never leaves a truncated file that would block the re-run (overwrite refusal, :47-57)."""

    def __init__(self, filename):
        try:
            import h5py
            self.path = filename
            self.tmp = filename + '.part'
            self.h5 = h5py.File(self.tmp, 'w')
            self.store = None
        except ImportError:
            self.h5 = None
            self.path = filename + '.npz'
            self.tmp = self.path + '.part'
            self.store = {}

    def put(self, path, data, compress=False, dtype=None):
        if self.h5 is not None:
            group, name = path.split('/')
            g = self.h5.require_group(group)
            kw = dict(chunks=True, compression='gzip') if compress and np.ndim(data) > 0 else {}
            if dtype is not None:
                kw['dtype'] = dtype
            g.create_dataset(name, data=data, **kw)
        else:
            arr = np.asarray(data) if dtype is None else np.asarray(data, dtype=dtype)
            if arr.dtype == object:
                arr = np.asarray(str(data))
            self.store[path] = arr

    def close(self):
        if self.h5 is not None:
            self.h5.close()
        else:
            # np.savez_compressed with a chosen zlib level (it is fixed at 6 there: 1.5x the time of
            # level 4 for 3 % smaller files on ADC noise)
            import zipfile
            with zipfile.ZipFile(self.tmp, 'w', zipfile.ZIP_DEFLATED, compresslevel=COMPRESS_LEVEL) as z:
                for key, arr in self.store.items():
                    with z.open(key + '.npy', 'w', force_zip64=True) as f:
                        np.lib.format.write_array(f, np.asanyarray(arr), allow_pickle=False)
        os.replace(self.tmp, self.path)
        return self.path


def output_exists(filename):
    return os.path.exists(filename) or os.path.exists(filename + '.npz')


SLAB_BYTES = 1 << 30       # pinned host memory per level: the cube crosses PCIe in slabs of this size


def produce_level(config, filename, n_events, seed, device, finish=None):
    """One (nsb, n_photon) level -> one file (reference produce_data.py:95-142).  ``config`` is the
    option dictionary: every entry goes into the ``config`` group, and it is the generator's keyword
    set (unknown keys are ignored there, as in the reference).  ``finish`` (optional) takes the
    zero-argument function that compresses and closes the file, so a caller can run it on another
    thread while the GPU generates the next level; the default runs it here.  The output file is
    only created once the level has been generated."""
    import torch
    from digicamtoy_b200 import NTraceGenerator
    log = logging
    kwargs = dict(config)
    kwargs.pop('seed', None)
    generator = NTraceGenerator(seed=seed, device=device, **kwargs)
    p = generator.params
    true_baseline, gen_seed = generator.true_baseline, generator.seed
    # replaces the loop :119-126; a level larger than one slab goes through one reused pinned slab
    slab_events = max(1, min(int(n_events), SLAB_BYTES // max(1, p.n_pixels * p.n_bins * 2)))
    pinned = ()
    if slab_events >= n_events:
        from digicamtoy_b200.tracegenerator import PINNED
        adc_count, cherenkov_time, cherenkov_photon = generator.generate_host(n_events, first_event=0, pool=PINNED)
        pinned = generator.last_host_tensors              # back to the pool once the file is written
    else:
        adc_count = np.empty((n_events, p.n_pixels, p.n_bins), dtype=np.uint16 if generator.saturate[0] >= 0 else np.int16)
        cherenkov_time = np.empty((n_events, p.n_pixels, p.n_pulses), dtype=np.float32)
        cherenkov_photon = np.empty_like(cherenkov_time)
        slab = torch.empty((slab_events, p.n_pixels, p.n_bins), dtype=torch.int16).pin_memory()
        for e0 in range(0, n_events, slab_events):
            n = min(slab_events, n_events - e0)
            a, t, q = generator.generate_host(n, first_event=e0, out=slab[:n])
            adc_count[e0:e0 + n], cherenkov_time[e0:e0 + n], cherenkov_photon[e0:e0 + n] = a, t, q
        del slab
    generator.close()                                                     # (the arrays keep their pinned buffers alive)

    def close():
        writer = Writer(filename)
        log.info('\t\t-|> Saving data to {}'.format(writer.path))
        for key, val in config.items():                                   # config group (:95-100)
            writer.put('config/' + key, val)
        writer.put('config/seed', np.uint64(gen_seed))
        writer.put('data/true_baseline', true_baseline, compress=True)
        writer.put('data/adc_count', adc_count, compress=True, dtype=np.uint16)
        writer.put('data/time', cherenkov_time.astype(np.float64), compress=True)
        writer.put('data/charge', cherenkov_photon.astype(np.float64), compress=True)
        writer.put('config/date', str(datetime.datetime.now()))
        path = writer.close()
        if pinned:
            from digicamtoy_b200.tracegenerator import PINNED
            PINNED.release(*pinned)
        log.info('\t\t-|> File saved! Size = {}'.format(natural_size(os.path.getsize(path))))
        return path

    if finish is not None:
        finish(close)
        return filename if _have_h5py() else filename + '.npz'
    return close()


def _have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except ImportError:
        return False


def main(argv=None):
    log = logging
    options, cli = load_options(argv)
    log.getLogger().setLevel(logging.DEBUG if options.debug else logging.INFO)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    device = cli['cli_device'] if cli['cli_device'] is not None else int(os.environ.get('LOCAL_RANK', '0'))

    # the reference refuses to overwrite (produce_data.py:47-57).  Under torchrun every rank checks
    # the files it will write itself: a rank that starts late must not trip over a file another
    # rank has already finished.
    for number, filename in enumerate(level_filenames(options)):
        if number % world == rank and output_exists(filename):
            log.error('File {} already exists'.format(filename))
            return 1

    log.debug(vars(options))
    log.info('\t\t-|> Create the Monte Carlo data set')

    options_generator = copy.copy(options)
    events_per_level = options_generator.events_per_level
    base_seed = cli['cli_seed']
    file_number = 0
    written = []
    for nsb_rate in options.nsb_rate:                                     # NSB-major numbering (:71-93)
        options_generator.nsb_rate = nsb_rate
        for n_photon in options.n_photon:
            options_generator.n_photon = n_photon
            my_turn = (file_number % world) == rank
            filename = os.path.join(options.output_directory, options.file_basename.format(file_number))
            if my_turn:
                written.append(produce_level(vars(options_generator), filename, events_per_level,
                                             None if base_seed is None else base_seed + file_number, device))
            file_number += 1
    return written


if __name__ == '__main__':
    out = main()
    sys.exit(out if isinstance(out, int) else 0)
