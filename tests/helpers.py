"""Shared helpers for the GPU parity tests (imports the oracle: test infrastructure only)."""
import glob
import json
import os

import numpy as np

from conftest import GOLDEN_DIR

GOLDEN_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    meta = json.loads(str(z['meta']))
    return z, meta


def pack_event_lists(events):
    """events: list of (nsb_time (P,M), nsb_amp (P,M), sig_time (P,K), sig_amp (P,K), noise (P,B)).
    Returns CSR offsets + flat arrays in [event, pixel] order."""
    offs = [0]
    t_all, a_all, st, sa, nz = [], [], [], [], []
    for nsb_t, nsb_a, s_t, s_a, noise in events:
        P, M = nsb_t.shape
        for p in range(P):
            t_all.append(nsb_t[p])
            a_all.append(nsb_a[p])
            offs.append(offs[-1] + M)
        st.append(s_t)
        sa.append(s_a)
        nz.append(noise)
    return (np.asarray(offs, dtype=np.int64), np.concatenate(t_all).astype(np.float64),
            np.concatenate(a_all).astype(np.float64), np.stack(st).astype(np.float64),
            np.stack(sa).astype(np.float64), np.stack(nz).astype(np.float64))


def replay_on_gpu(handle, precision, events, n_pixels, n_bins):
    import torch
    offs, t, a, st, sa, nz = pack_event_lists(events)
    dev = torch.device('cuda', handle.device)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    offs_d, t_d, a_d, st_d, sa_d, nz_d = d(offs), d(t), d(a), d(st), d(sa), d(nz)
    if t_d.numel() == 0:
        t_d = torch.zeros(1, dtype=torch.float64, device=dev)
        a_d = torch.zeros(1, dtype=torch.float64, device=dev)
    n_ev = len(events)
    adc = torch.empty((n_ev, n_pixels, n_bins), dtype=torch.int16, device=dev)
    clamped = torch.zeros(1, dtype=torch.int64, device=dev)
    handle.replay(precision, n_ev, offs_d.data_ptr(), t_d.data_ptr(), a_d.data_ptr(), st_d.data_ptr(),
                  sa_d.data_ptr(), nz_d.data_ptr(), adc.data_ptr(), clamped.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return adc.cpu().numpy(), int(clamped.item())
