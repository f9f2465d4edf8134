"""Host side of the 12-bit transport of dct_generate_host (csrc/host_transport.h): the widening of a
packed stream into int16 -- scalar loop, AVX2 path and the thread pool -- against a NumPy packing of
the same samples.  No GPU involved."""
import numpy as np
import pytest

from digicamtoy_b200 import native


def pack12(samples):
    """NumPy reference of the stream the kernels write: sample s in bits [12 s, 12 s + 12), little endian."""
    s = np.asarray(samples, dtype=np.uint32)
    n = len(s)
    if n % 2:
        s = np.append(s, 0)
    a, b = s[0::2], s[1::2]
    out = np.empty((len(a), 3), dtype=np.uint8)
    out[:, 0] = a & 0xff
    out[:, 1] = (a >> 8) | ((b & 0xf) << 4)
    out[:, 2] = b >> 4
    out = out.reshape(-1)
    return out[:(3 * n + 1) // 2]


@pytest.mark.parametrize('n', [0, 1, 2, 3, 15, 16, 17, 31, 32, 33, 47, 48, 49, 1000, 64801, (1 << 20) + 5, 3 * (1 << 20) + 1])
@pytest.mark.parametrize('threads,offset', [(0, 0), (1, 0), (3, 0), (1, 2), (1, 1), (2, 7)])
def test_unpack12_equals_the_numpy_packing(n, threads, offset):
    rs = np.random.RandomState(n % 9973 + threads)
    x = rs.randint(0, 4096, size=n).astype(np.int16)
    if n > 4:
        x[:4] = [0, 4095, 1, 4094]
    packed = pack12(x)
    assert len(packed) == (3 * n + 1) // 2
    # guard bytes behind the stream: the widening must not depend on what follows it
    padded = np.concatenate([packed, rs.randint(0, 256, size=64).astype(np.uint8)])
    out = native.test_unpack12(padded, n, threads, offset)      # offsets: streaming stores after a peel / unaligned stores
    assert out.dtype == np.int16 and np.array_equal(out, x)


def test_unpack12_rejects_null_pointers():
    lib = native.load_library()
    assert lib.dct_test_unpack12(None, None, 4, 0) != 0
    assert b'NULL' in lib.dct_last_error()


@pytest.mark.parametrize('unit,n', [(219, 16384), (219, 1), (219, 219), (219, 220), (219, 8 * 219), (219, 8 * 219 + 1),
                                    (129, 1500), (1, 7), (1, 100), (10168, 53), (73, 10 ** 6)])
def test_chunk_schedule_of_the_host_pipeline(unit, n):
    """dct_generate_host walks a request in chunks of 1, 2, 4 ... 4, rest, 2, 1 units (all of one unit up to
    eight units): contiguous, no empty chunk, nothing larger than the four-unit device buffers."""
    ends = native.test_chunk_schedule(unit, n).astype(np.int64)
    sizes = np.diff(np.concatenate([[0], ends]))
    assert ends[-1] == n and (sizes > 0).all() and sizes.max() <= 4 * unit
    units = -(-n // unit)
    if units <= 8:
        assert len(sizes) == units and (sizes[:-1] == unit).all()
    else:
        assert list(sizes[:2]) == [unit, 2 * unit] and sizes[-1] <= unit and sizes[-2] <= 2 * unit
        assert (sizes[2:-3] == 4 * unit).all()                      # the middle: full buffers, then one rest


def test_chunk_schedule_rejects_bad_arguments():
    lib = native.load_library()
    assert lib.dct_test_chunk_schedule(0, 10, None, 0) < 0


def test_iterator_batch_sizes_ramp_up_on_long_runs():
    """NTraceGenerator._batch_size_at (host logic, no device): the first batches of a long run are small
    (32 MiB), then double up to the 128 MiB steady state; a run with a known end never asks for more
    than it has left."""
    from digicamtoy_b200.tracegenerator import NTraceGenerator
    g = object.__new__(NTraceGenerator)
    g.batch_events, g._first_batch, g.n_events = 1035, 258, None
    starts, at = [], 0
    for _ in range(7):
        starts.append((at, g._batch_size_at(at)))
        at += starts[-1][1]
    assert starts == [(0, 258), (258, 258), (516, 516), (1032, 1032), (2064, 1035), (3099, 1035), (4134, 1035)]
    g.n_events = 3000
    assert g._batch_size_at(2064) == 936 and g._batch_size_at(3000) == 0
    g.batch_events, g._first_batch, g.n_events = 10, 10, 53          # explicit batch size: no ramp
    assert [g._batch_size_at(k) for k in (0, 10, 50)] == [10, 10, 3]
