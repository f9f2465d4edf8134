"""T5: the N>1 plumbing on CPU (gloo, world_size 2): event shards are disjoint and complete, and
the statistics all-reduce of a sharded run equals the single-process sum."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_events, out_dir):
    sys.path.insert(0, ROOT)
    from digicamtoy_b200 import workloads
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    first, count = workloads.shard_events(n_events, rank, world)
    # stand-in for the per-rank statistics buffers (CubeStats._f / ._i): a histogram of the global
    # event indices this rank owns plus a power sum
    hist = torch.zeros(n_events, dtype=torch.int64)
    hist[first:first + count] += 1
    sums = torch.tensor([float(sum(range(first, first + count))), float(count)], dtype=torch.float64)
    dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)              # max-over-ranks timing reduction
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.savez(os.path.join(out_dir, 'r.npz'), hist=hist.numpy(), sums=sums.numpy(), t=t.numpy())
    dist.destroy_process_group()


def test_sharded_statistics_reduce_to_single_process_result(tmp_path):
    n_events, world = 101, 2
    mp.spawn(_worker, args=(world, _free_port(), n_events, str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / 'r.npz')
    assert np.array_equal(z['hist'], np.ones(n_events, dtype=np.int64))     # every event exactly once
    assert z['sums'][0] == sum(range(n_events)) and z['sums'][1] == n_events
    assert z['t'][0] == 2.0


def _stats_worker(rank, world, port, n_events, out_dir):
    """The product's own accumulator and its all_reduce (CubeStats), on CPU tensors over gloo: every
    rank adds the statistics of its shard of one synthetic cube; rank 0 saves the reduced result."""
    sys.path.insert(0, ROOT)
    from digicamtoy_b200 import workloads
    from digicamtoy_b200.stats import CubeStats
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    cube = _synthetic_cube(n_events)
    first, count = workloads.shard_events(n_events, rank, world)
    st = CubeStats.from_shapes(7, 10, saturate=(0, 4095), baseline=200.4, charge_lo=-100, n_charge_bins=400)
    if count:
        st.accumulate_numpy(cube[first:first + count])
    st.all_reduce()
    if rank == 0:
        r = st.result()
        np.savez(os.path.join(out_dir, 's.npz'), moments_nsb=np.array(st.moments_nsb()), **r)
    dist.destroy_process_group()


def _synthetic_cube(n_events):
    rs = np.random.RandomState(5)
    c = rs.normal(200, 3, size=(n_events, 7, 10)).round().astype(np.int64)
    c[3, 2, :] = 4095                                                    # a saturated trace: charge beyond the last bin
    c[4, 1, :] = 0
    return c


def test_cube_stats_all_reduce_over_gloo_equals_the_unsharded_statistics(tmp_path):
    from digicamtoy_b200.stats import CubeStats
    n_events, world = 23, 2
    mp.spawn(_stats_worker, args=(world, _free_port(), n_events, str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / 's.npz')
    whole = CubeStats.from_shapes(7, 10, saturate=(0, 4095), baseline=200.4, charge_lo=-100, n_charge_bins=400)
    whole.accumulate_numpy(_synthetic_cube(n_events))
    r = whole.result()
    assert int(z['n_events']) == n_events == r['n_events']
    for key in ('adc_hist', 'charge_hist', 'bin_sum', 'bin_sumsq', 'pixel_sum', 'pixel_sumsq'):
        assert np.array_equal(z[key], r[key]), key
    assert np.allclose(z['moments'], r['moments'], rtol=1e-14)
    c = _synthetic_cube(n_events).astype(float)
    assert z['moments_nsb'][0] == c.mean() and abs(z['moments_nsb'][2] - c.std()) < 1e-9
    assert r['charge_hist'][-1] == 1 and r['charge_hist'][0] == 1      # the two rail traces, clamped


def test_cube_stats_all_reduce_is_noop_without_process_group():
    from digicamtoy_b200.stats import moments_from_power_sums
    x = np.random.RandomState(0).normal(300, 10, 5000).round()
    mean, mean_err, std, std_err = moments_from_power_sums(
        [x.sum(), (x ** 2).sum(), (x ** 3).sum(), (x ** 4).sum()], x.size)
    from scipy.stats import moment
    assert abs(mean - x.mean()) < 1e-9 and abs(std - x.std()) < 1e-6
    m4 = moment(x, moment=4)
    want = 0.5 * np.sqrt(1. / x.size * (m4 - (x.size - 3) / (x.size - 1) * x.std() ** 4)) / x.std()
    assert abs(std_err - want) < 1e-6 * want + 1e-9
    assert abs(mean_err - x.std() / np.sqrt(x.size)) < 1e-9
