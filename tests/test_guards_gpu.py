"""Out-of-bounds canaries (compute-sanitizer is closed on this pool): every kernel writes into the
middle of a sentinel-filled buffer; the guard regions must come back untouched.  Covers the ragged
last CTA, odd bin counts, 16-byte-misaligned destinations and every kernel variant."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    dict(n_pixels=37, nsb_rate=0.5, n_photon=8., jitter=0.2),                                   # sweep
    dict(n_pixels=130, nsb_rate=0.5, n_photon=8., kernel='scatter'),                            # scatter
    dict(n_pixels=5, time_end=368, nsb_rate=0.003, n_photon=2, jitter=0.3),                     # 92 bins
    dict(n_pixels=9, time_end=199, nsb_rate=0.3),                                               # 50 bins, trace count 9k
    dict(n_pixels=7, time_start=0, time_end=102, time_sampling=6, nsb_rate=0.3),               # 17 bins (odd)
    dict(n_pixels=6, nsb_rate=0.3, sub_binning=1, n_photon=4),
    dict(n_pixels=6, nsb_rate=0.5, pulse_shape_file='utils/pulse_SST-1M_AfterPreampLowGain.dat'),
]


@pytest.mark.parametrize('kw', CASES)
@pytest.mark.parametrize('shift', [0, 8, 3])          # elements: 0 / 16-byte aligned / misaligned
def test_generate_never_writes_outside_its_cube(kw, shift):
    import torch
    from digicamtoy_b200 import NTraceGenerator
    g = NTraceGenerator(seed=5, **kw)
    P, B, K = g.params.n_pixels, g.params.n_bins, g.params.n_pulses
    for n in (1, 5):
        pad = 4096
        buf = torch.full((pad + shift + n * P * B + pad,), -12345, dtype=torch.int16, device='cuda')
        sig = torch.full((2, 64 + n * P * K + 64), -7.0, dtype=torch.float32, device='cuda')
        g._handle.generate(g.seed, 0, n, buf.data_ptr() + 2 * (pad + shift), sig[0].data_ptr() + 4 * 64,
                           sig[1].data_ptr() + 4 * 64, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        b = buf.cpu().numpy()
        assert (b[:pad + shift] == -12345).all() and (b[pad + shift + n * P * B:] == -12345).all()
        body = b[pad + shift:pad + shift + n * P * B]
        assert body.min() >= 0 and body.max() <= 4095
        s = sig.cpu().numpy()
        assert (s[:, :64] == -7.0).all() and (s[:, 64 + n * P * K:] == -7.0).all()
        ref = g.generate(n, first_event=0)
        torch.cuda.synchronize()
        assert np.array_equal(body.reshape(n, P, B), ref.cpu().numpy())      # same bits on every store path


@pytest.mark.parametrize('kw', CASES[:5])
def test_statistics_kernels_stay_inside_their_buffers(kw):
    import torch
    from digicamtoy_b200 import NTraceGenerator
    g = NTraceGenerator(seed=5, **kw)
    P, B = g.params.n_pixels, g.params.n_bins
    cube = g.generate(11)
    G = 32
    nh, nc = 4096, 777
    f = torch.full((G + 2 * B + 2 * P + 4 + G,), -1.0, dtype=torch.float64, device='cuda')
    i = torch.full((G + nh + nc + G,), -1, dtype=torch.int64, device='cuda')
    f[G:-G] = 0
    i[G:-G] = 0
    fp, ip = f.data_ptr() + 8 * G, i.data_ptr() + 8 * G
    g._handle.stats_accumulate(cube.data_ptr(), 11, adc_hist=ip, bin_sum=fp, bin_sumsq=fp + 8 * B,
                               pixel_sum=fp + 16 * B, pixel_sumsq=fp + 16 * B + 8 * P,
                               charge_hist=(ip + 8 * nh) if B <= 128 else None, charge_lo=-300,
                               n_charge_bins=nc, moments=fp + 16 * B + 16 * P,
                               stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    fh, ih = f.cpu().numpy(), i.cpu().numpy()
    assert (fh[:G] == -1).all() and (fh[-G:] == -1).all() and (ih[:G] == -1).all() and (ih[-G:] == -1).all()
    c = cube.cpu().numpy().astype(np.int64)
    assert ih[G:G + nh].sum() == c.size and ih[G + nh:G + nh + nc].sum() == 11 * P
    assert np.allclose(fh[G:G + B], c.sum(axis=(0, 1)))
    assert np.allclose(fh[G + 2 * B:G + 2 * B + P], c.sum(axis=(0, 2)))
    assert np.allclose(fh[-G - 4:-G], [(c.astype(float) ** k).sum() for k in (1, 2, 3, 4)], rtol=1e-12)


def test_replay_and_trace_pe_guards():
    import torch
    from digicamtoy_b200 import NTraceGenerator
    g = NTraceGenerator(seed=9, n_pixels=11, nsb_rate=0.4, n_photon=5.)
    P, B = 11, 50
    n = 3
    cnt = torch.full((16 + n * P + 16,), -5, dtype=torch.int32, device='cuda')
    g._handle.trace_pe(g.seed, 0, n, 0, cnt.data_ptr() + 4 * 16)
    torch.cuda.synchronize()
    m = int(cnt[16:-16].max())
    t = torch.full((8 + n * P * m + 8,), -9.0, dtype=torch.float64, device='cuda')
    a = torch.full((8 + n * P * m + 8,), -9.0, dtype=torch.float32, device='cuda')
    z = torch.full((8 + n * P * B + 8,), -9.0, dtype=torch.float32, device='cuda')
    g._handle.trace_pe(g.seed, 0, n, m, cnt.data_ptr() + 4 * 16, t.data_ptr() + 64, a.data_ptr() + 32,
                       z.data_ptr() + 32)
    torch.cuda.synchronize()
    for buf, g0 in ((cnt, 16), (t, 8), (a, 8), (z, 8)):
        h = buf.cpu().numpy()
        assert (h[:g0] == h[0]).all() and (h[-g0:] == h[0]).all() and h[0] < 0
    assert (cnt[16:-16] >= 0).all()
