"""Host logic of the produce_data drop-in and of event sharding (no GPU needed)."""
import os

import numpy as np
import pytest

import produce_data
from digicamtoy_b200 import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_yaml_options_merge_like_reference(tmp_path):
    opts, cli = produce_data.load_options(['-y', os.path.join(ROOT, 'config_files', 'dark.yml'),
                                           '--output_directory', str(tmp_path), '--events_per_level', '7'])
    assert opts.n_pixels == 1296 and opts.time_end == 368 and opts.nsb_rate == [0.003]
    assert opts.n_photon == [0., 1, 2, 3] and opts.events_per_level == 7
    assert opts.output_directory == str(tmp_path) and opts.debug is False
    names = produce_data.level_filenames(opts)
    assert [os.path.basename(n) for n in names] == ['dark_digicamtoy_%d.hdf5' % i for i in range(4)]
    assert not any(k.startswith('cli_') for k in vars(opts))


def test_ac_dc_levels_are_nsb_major(tmp_path):
    opts, _ = produce_data.load_options(['-y', os.path.join(ROOT, 'config_files', 'ac_dc.yml'),
                                         '--output_directory', str(tmp_path)])
    assert len(produce_data.level_filenames(opts)) == 24            # 6 NSB levels x 4 n_photon levels
    assert opts.events_per_level == 10 and opts.sigma_e == 1.25 and opts.baseline == 280


def test_refuses_to_overwrite(tmp_path):
    for suffix in ('', '.npz'):
        d = tmp_path / ('x' + suffix.strip('.'))
        d.mkdir()
        (d / ('dark_digicamtoy_2.hdf5' + suffix)).write_bytes(b'')
        rc = produce_data.main(['-y', os.path.join(ROOT, 'config_files', 'dark.yml'),
                                '--output_directory', str(d)])
        assert rc == 1


def test_npz_writer_keeps_dataset_paths(tmp_path):
    try:
        import h5py  # noqa: F401
        pytest.skip('h5py present: HDF5 path is used')
    except ImportError:
        pass
    w = produce_data.Writer(str(tmp_path / 'f.hdf5'))
    w.put('config/n_photon', [0., 1, 2])
    w.put('config/pulse_shape_file', 'utils/x.dat')
    w.put('data/adc_count', np.arange(24).reshape(2, 3, 4), compress=True, dtype=np.uint16)
    path = w.close()
    assert path.endswith('f.hdf5.npz')
    z = np.load(path)
    assert z['data/adc_count'].dtype == np.uint16 and z['data/adc_count'].shape == (2, 3, 4)
    assert str(z['config/pulse_shape_file']) == 'utils/x.dat'


def test_natural_size():
    assert produce_data.natural_size(10) == '10 Bytes'
    assert produce_data.natural_size(5 * 1024 ** 2) == '5.0 MiB'


def test_event_shards_cover_range_without_overlap():
    for n, w in ((10, 1), (10, 3), (7, 8), (10 ** 7, 8), (0, 4)):
        pieces = [workloads.shard_events(n, r, w) for r in range(w)]
        assert sum(c for _, c in pieces) == n
        cursor = 0
        for first, count in pieces:
            assert first == cursor
            cursor += count


def test_named_workloads():
    c4 = workloads.c4_full_camera()
    assert c4['n_pixels'] == 1296 and c4['nsb_rate'] == 1.0 and len(c4['n_photon']) == 1296
    assert c4['n_photon'][0] == [1.0] and abs(c4['n_photon'][-1][0] - 10 ** 3.5) < 1e-9
    d = workloads.c1_dark(2)
    assert d['time_end'] == 368 and d['nsb_rate'] == 0.003 and d['n_photon'] == 2
    a = workloads.c2_ac_dc(0.66, 100)
    assert a['nsb_rate'] == 0.66 and a['n_photon'] == 100 and a['sigma_e'] == 1.25


# ------------------------------------------------------------------ f-4: the commissioning grid launcher
def test_batch_grid_follows_the_reference_job_array():
    import batch_produce as bp
    jobs = bp.job_grid((0, 499), (0, 0), (0, 9))                 # jobs.sh:3-16
    assert len(jobs) == 5000 and jobs[0] == (0, 0, 0) and jobs[1] == (0, 0, 1) and jobs[10] == (1, 0, 0)
    assert jobs[-1] == (499, 0, 9)
    parts = [bp.deal(jobs, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == sorted(jobs) and {len(p) for p in parts} == {625}
    seeds = {bp.job_seed(100, 30, 10, *j) for j in jobs}
    assert len(seeds) == len(jobs) and bp.job_seed(None, 30, 10, 1, 2, 3) is None


def test_batch_configs_equal_what_generate_py_dumps(tmp_path):
    """Same keys and values as reference config_files/generate.py:8-71 (restated here)."""
    import yaml
    import batch_produce as bp
    paths = bp.main(['--write_configs', str(tmp_path), '--ac', '3:4', '--file_id', '0:1'])
    assert [os.path.basename(p) for p in paths] == ['ac_3_dc_0_id_0.yml', 'ac_3_dc_0_id_1.yml',
                                                    'ac_4_dc_0_id_0.yml', 'ac_4_dc_0_id_1.yml']
    with open(paths[2]) as f:
        got = yaml.safe_load(f)
    P = 1296
    want = dict(output_directory='/sst1m/MC/digicamtoy/', events_per_level=1000, time_start=0, time_end=368,
                time_sampling=4, n_pixels=P, gain_nsb=True, poisson=True,
                pulse_shape_file='utils/pulse_SST-1M_pixel_0.dat', sigma_e=[0.8] * P, sigma_1=[0.8] * P,
                gain=[5.8] * P, baseline=[200] * P, time_signal=[20] * P, jitter=[1.155] * P,
                crosstalk=[0.08] * P, nsb_rate=[0.003], n_photon=[4.0], file_basename='ac_4_dc_0_id_0.hdf5')
    assert got == want
    n_photons, nsb_rates = bp.light_levels()
    assert len(n_photons) == 500 and len(nsb_rates) == 30 and nsb_rates[-1] == 2.0 and n_photons[-1] == 499


def test_batch_dry_run_and_overwrite_refusal(tmp_path, capsys):
    import batch_produce as bp
    files = bp.main(['--dry_run', '--ac', '0:1', '--file_id', '0:2', '--output_directory', str(tmp_path)])
    assert len(files) == 6 and files[0].endswith('ac_0_dc_0_id_0.hdf5')
    assert len(capsys.readouterr().out.strip().splitlines()) == 6
    (tmp_path / 'ac_1_dc_0_id_2.hdf5.npz').write_bytes(b'')
    assert bp.main(['--ac', '0:1', '--file_id', '0:2', '--output_directory', str(tmp_path)]) == 1
    with pytest.raises(ValueError):
        bp.main(['--dry_run', '--ac', '0:500'])


def test_hdf5_branch_of_the_writer_with_a_stub_h5py(tmp_path, monkeypatch):
    """h5py is not in this image; a recording stub checks the calls of the HDF5 branch
    (groups, gzip-chunked datasets, dtype) -- reference produce_data.py:95-142."""
    import sys
    import types
    calls = []

    class Group:
        def __init__(self, name):
            self.name = name

        def create_dataset(self, name, data=None, **kw):
            calls.append((self.name, name, np.asarray(data).shape if not isinstance(data, str) else 'str', kw))

    class File:
        def __init__(self, filename, mode):
            assert mode == 'w'
            self.filename, self.groups, self.closed = filename, {}, False

        def require_group(self, name):
            return self.groups.setdefault(name, Group(name))

        def close(self):
            self.closed = True
            open(self.filename, 'wb').close()               # the real h5py leaves a file behind

    stub = types.ModuleType('h5py')
    stub.File = File
    monkeypatch.setitem(sys.modules, 'h5py', stub)
    w = produce_data.Writer(str(tmp_path / 'f.hdf5'))
    assert w.h5 is not None and w.path.endswith('f.hdf5')
    w.put('config/n_pixels', 1296)
    w.put('config/pulse_shape_file', 'utils/x.dat')
    w.put('data/adc_count', np.zeros((2, 3, 4), dtype=np.int16), compress=True, dtype=np.uint16)
    w.put('data/true_baseline', np.zeros(3), compress=True)
    assert w.h5.filename.endswith('f.hdf5.part')            # written under a temporary name ...
    assert w.close().endswith('f.hdf5') and w.h5.closed
    assert (tmp_path / 'f.hdf5').exists() and not (tmp_path / 'f.hdf5.part').exists()      # ... renamed when complete
    assert calls[0] == ('config', 'n_pixels', (), {})
    assert calls[1] == ('config', 'pulse_shape_file', 'str', {})
    assert calls[2] == ('data', 'adc_count', (2, 3, 4), dict(chunks=True, compression='gzip', dtype=np.uint16))
    assert calls[3] == ('data', 'true_baseline', (3,), dict(chunks=True, compression='gzip'))


def test_batch_ranks_partition_the_grid(tmp_path, monkeypatch, capsys):
    """Under torchrun every rank walks its own share: together exactly the grid, no file twice."""
    import batch_produce as bp
    seen = []
    for rank in range(3):
        monkeypatch.setenv('RANK', str(rank))
        monkeypatch.setenv('WORLD_SIZE', '3')
        seen.append(bp.main(['--dry_run', '--ac', '5:8', '--file_id', '0:4', '--output_directory', str(tmp_path)]))
    capsys.readouterr()
    flat = [f for part in seen for f in part]
    assert len(flat) == 20 and len(set(flat)) == 20
    assert {len(part) for part in seen} == {6, 7}
    assert os.path.basename(seen[1][0]) == 'ac_5_dc_0_id_1.hdf5'
