"""f-1 / (e): the produce_data drop-in end to end, and sharded == unsharded on one GPU."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path):
    if path.endswith('.npz'):
        z = np.load(path)
        return {k: z[k] for k in z.files}
    import h5py
    out = {}
    with h5py.File(path, 'r') as f:
        for g in ('config', 'data'):
            for k in f[g]:
                out[g + '/' + k] = f[g][k][()]
    return out


def test_produce_data_dark_yaml(tmp_path):
    import produce_data
    written = produce_data.main(['-y', os.path.join(ROOT, 'config_files', 'dark.yml'),
                                 '--output_directory', str(tmp_path), '--events_per_level', '12',
                                 '--seed', '5'])
    assert len(written) == 4                                       # n_photon [0,1,2,3] x nsb [0.003]
    for i, path in enumerate(written):
        assert os.path.basename(path).startswith('dark_digicamtoy_%d.hdf5' % i)
        d = _load(path)
        assert d['data/adc_count'].shape == (12, 1296, 92) and d['data/adc_count'].dtype == np.uint16
        assert d['data/time'].shape == (12, 1296, 1) and d['data/charge'].shape == (12, 1296, 1)
        assert d['data/true_baseline'].shape == (1296,)
        assert float(d['config/n_photon']) == [0., 1, 2, 3][i] and float(d['config/nsb_rate']) == 0.003
        assert 'config/date' in d and 'config/pulse_shape_file' in d
        # pedestal + the pulse: i pe x 1/(1-xt) x gain_eff x sum_b h(t_b) / 92 bins ~ 0.307 LSB per pe
        shift = d['data/adc_count'].mean() - d['data/true_baseline'].mean()
        assert abs(shift - 0.307 * i) < 0.12, shift
        if i == 0:
            assert not d['data/charge'].any()
        else:
            assert abs(d['data/charge'].mean() - i / (1 - 0.08)) < 0.1
    # second run refuses to overwrite
    assert produce_data.main(['-y', os.path.join(ROOT, 'config_files', 'dark.yml'),
                              '--output_directory', str(tmp_path), '--events_per_level', '12']) == 1


def test_sharded_run_equals_single_run():
    """World size 4 emulated on one GPU: each 'rank' generates its contiguous global-event range
    with the same seed; cubes concatenate to the single-GPU cube and statistics add up."""
    import torch
    from digicamtoy_b200 import NTraceGenerator, workloads
    from digicamtoy_b200.stats import CubeStats
    kw = dict(n_pixels=96, nsb_rate=0.66, n_photon=10., time_start=0, time_end=200, time_sampling=4,
              sigma_e=1.25, gain=5, baseline=280, jitter=0.2)
    n_events, world = 101, 4
    single = NTraceGenerator(seed=42, **kw)
    whole = single.generate(n_events)
    st_whole = CubeStats(single, charge_lo=-1000, n_charge_bins=8000)
    st_whole.accumulate(whole)
    want = st_whole.result()
    parts, acc_f, acc_i = [], None, None
    for rank in range(world):
        g = NTraceGenerator(seed=42, **kw)
        first, count = workloads.shard_events(n_events, rank, world)
        cube = g.generate(count, first_event=first)
        st = CubeStats(g, charge_lo=-1000, n_charge_bins=8000)
        st.accumulate(cube)
        acc_f = st._f.clone() if acc_f is None else acc_f + st._f       # what all_reduce(SUM) does
        acc_i = st._i.clone() if acc_i is None else acc_i + st._i
        parts.append(cube)
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(parts), whole)
    assert np.array_equal(acc_i.cpu().numpy(), st_whole._i.cpu().numpy())
    assert np.allclose(acc_f.cpu().numpy(), st_whole._f.cpu().numpy(), rtol=1e-12)
    assert want['n_events'] == n_events


def test_batch_launcher_writes_the_commissioning_grid(tmp_path, monkeypatch):
    """f-4: (n_photon level, NSB level, file id) grid in one process; a file equals what
    produce_data writes from the YAML that generate.py would have dumped for that job."""
    import batch_produce as bp
    import produce_data
    out = tmp_path / 'mc'
    common = ['--n_pixels', '48', '--events_per_level', '9', '--ac', '2:3', '--dc', '0:1', '--file_id', '0:1']
    written = bp.main(common + ['--output_directory', str(out), '--seed', '1000', '--writers', '2',
                                '--max_in_flight', '2'])
    assert len(written) == 8
    names = [os.path.basename(p).replace('.npz', '') for p in written]
    assert names[:3] == ['ac_2_dc_0_id_0.hdf5', 'ac_2_dc_0_id_1.hdf5', 'ac_2_dc_1_id_0.hdf5']
    cubes = {}
    for p, n in zip(written, names):
        d = _load(p)
        assert d['data/adc_count'].shape == (9, 48, 92) and d['data/adc_count'].dtype == np.uint16
        assert d['data/time'].shape == (9, 48, 1)
        i, j = int(n.split('_')[1]), int(n.split('_')[3])
        assert float(d['config/n_photon']) == float(i)
        assert float(d['config/nsb_rate']) == pytest.approx(np.linspace(0.003, 2, 30)[j])
        assert float(np.ravel(d['config/jitter'])[0]) == 1.155 and int(d['config/time_end']) == 368
        cubes[n] = d
    # different files of one level are independent samples; seeds follow (i, j, k)
    assert not np.array_equal(cubes['ac_2_dc_0_id_0.hdf5']['data/adc_count'],
                              cubes['ac_2_dc_0_id_1.hdf5']['data/adc_count'])
    n_dc, n_ids = 30, 2
    assert int(cubes['ac_3_dc_1_id_1.hdf5']['config/seed']) == 1000 + (3 * n_dc + 1) * n_ids + 1
    # the NSB level shows: dc_1 (72 MHz) sits above dc_0 (3 MHz)
    assert cubes['ac_2_dc_1_id_0.hdf5']['data/adc_count'].mean() > cubes['ac_2_dc_0_id_0.hdf5']['data/adc_count'].mean() + 3
    # the YAML route (generate.py step + produce_data.py -y) gives the same file content
    cfg_dir = tmp_path / 'commissioning'
    bp.main(common + ['--write_configs', str(cfg_dir), '--output_directory', str(tmp_path / 'mc2')])
    seed = 1000 + (3 * n_dc + 1) * n_ids + 1
    os.makedirs(str(tmp_path / 'mc2'))
    (path,) = produce_data.main(['-y', str(cfg_dir / 'ac_3_dc_1_id_1.yml'), '--seed', str(seed)])
    d2 = _load(path)
    assert np.array_equal(d2['data/adc_count'], cubes['ac_3_dc_1_id_1.hdf5']['data/adc_count'])
    assert np.array_equal(d2['data/charge'], cubes['ac_3_dc_1_id_1.hdf5']['data/charge'])
    # a rank of a 4-process launch only takes its share, and nothing is overwritten
    monkeypatch.setenv('RANK', '1')
    monkeypatch.setenv('WORLD_SIZE', '4')
    share = bp.main(common + ['--dry_run', '--output_directory', str(out)])
    assert [os.path.basename(s) for s in share] == ['ac_2_dc_0_id_1.hdf5', 'ac_3_dc_0_id_1.hdf5']
    assert bp.main(common + ['--output_directory', str(out)]) == 1
