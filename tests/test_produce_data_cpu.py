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
