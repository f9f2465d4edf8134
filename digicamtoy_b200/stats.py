"""On-device validation statistics of ADC cubes, with an NCCL/Gloo all-reduce for sharded runs.

The quantities are the ones the reference's analysis scripts compute on the host from the whole
cube: mean / std / their errors over waveforms x bins (reference ``nsb_baseline.py:21-32``), the
ADC spectrum (``digicamtoy/test/test.py:40-47``), and the integrated-charge spectrum.
"""
from __future__ import annotations

import numpy as np


class CubeStats:
    """Accumulates K3 statistics of cubes produced by one generator (same P, B, baseline)."""

    def __init__(self, generator, charge_lo=-2048, n_charge_bins=16384):
        import torch
        p = generator.params
        self._setup(torch, p.n_pixels, p.n_bins, generator.saturate, np.asarray(p.baseline, dtype=np.float64),
                    charge_lo, n_charge_bins, torch.device('cuda', generator.device))
        self.gen = generator

    @classmethod
    def from_shapes(cls, n_pixels, n_bins, saturate=(0, 4095), baseline=0.0, charge_lo=-2048, n_charge_bins=16384,
                    device='cpu'):
        """The accumulator without a generator: same buffers, same all-reduce, same results.  Cubes are
        added with ``accumulate_numpy`` (host arithmetic) -- for merging statistics files and for
        exercising the multi-rank reduction on machines without a GPU (gloo)."""
        import torch
        self = cls.__new__(cls)
        self._setup(torch, int(n_pixels), int(n_bins), tuple(saturate),
                    np.broadcast_to(np.asarray(baseline, dtype=np.float64), (int(n_pixels),)).copy(),
                    charge_lo, n_charge_bins, torch.device(device))
        self.gen = None
        return self

    def _setup(self, torch, n_pixels, n_bins, saturate, baseline, charge_lo, n_charge_bins, dev):
        self._torch = torch
        self.n_pixels, self.n_bins, self.saturate, self._baseline = n_pixels, n_bins, saturate, baseline
        lo, hi = saturate
        self.charge_lo, self.n_charge_bins = int(charge_lo), int(n_charge_bins)
        # one flat float64 + one flat int64 buffer: a sharded run all-reduces exactly two tensors
        self._f = torch.zeros(2 * n_bins + 2 * n_pixels + 4, dtype=torch.float64, device=dev)
        self._i = torch.zeros((hi - lo + 1) + self.n_charge_bins + 1, dtype=torch.int64, device=dev)
        self._slices_f = {}
        o = 0
        for name, n in (('bin_sum', n_bins), ('bin_sumsq', n_bins), ('pixel_sum', n_pixels),
                        ('pixel_sumsq', n_pixels), ('moments', 4)):
            self._slices_f[name] = slice(o, o + n)
            o += n
        self._slices_i = {'adc_hist': slice(0, hi - lo + 1),
                          'charge_hist': slice(hi - lo + 1, hi - lo + 1 + self.n_charge_bins),
                          'n_events': slice(hi - lo + 1 + self.n_charge_bins, hi - lo + 2 + self.n_charge_bins)}
        self.with_charge = n_bins <= 128

    def accumulate_numpy(self, cube):
        """Host-side twin of ``accumulate`` (NumPy arithmetic on a ``[n, P, B]`` integer array): the
        same quantities into the same buffers.  Not used by the generation path."""
        torch = self._torch
        c = np.asarray(cube).astype(np.int64)
        assert c.ndim == 3 and c.shape[1:] == (self.n_pixels, self.n_bins)
        lo, hi = self.saturate
        f = np.zeros(self._f.numel()); i = np.zeros(self._i.numel(), dtype=np.int64)
        i[self._slices_i['adc_hist']] = np.bincount((c - lo).ravel(), minlength=hi - lo + 1)
        f[self._slices_f['bin_sum']] = c.sum(axis=(0, 1)); f[self._slices_f['bin_sumsq']] = (c ** 2).sum(axis=(0, 1))
        f[self._slices_f['pixel_sum']] = c.sum(axis=(0, 2)); f[self._slices_f['pixel_sumsq']] = (c ** 2).sum(axis=(0, 2))
        x = c.ravel().astype(np.float64)
        f[self._slices_f['moments']] = [x.sum(), (x ** 2).sum(), (x ** 3).sum(), (x ** 4).sum()]
        if self.with_charge:
            base = np.rint(self._baseline.astype(np.float32)).astype(np.int64)
            charge = c.sum(axis=2) - self.n_bins * base[None, :] - self.charge_lo
            i[self._slices_i['charge_hist']] = np.bincount(np.clip(charge.ravel(), 0, self.n_charge_bins - 1),
                                                           minlength=self.n_charge_bins)
        i[self._slices_i['n_events']] = c.shape[0]
        self._f += torch.from_numpy(f).to(self._f.device)
        self._i += torch.from_numpy(i).to(self._i.device)

    def _ptr(self, buf, sl):
        return buf.data_ptr() + sl.start * buf.element_size()

    def accumulate(self, adc, stream=None):
        torch = self._torch
        if self.gen is None:
            raise RuntimeError('this accumulator has no generator: use accumulate_numpy')
        p = self.gen.params
        if adc.dtype != torch.int16 or not adc.is_contiguous() or not adc.is_cuda:
            raise ValueError('adc must be a contiguous int16 CUDA tensor')
        n_events = adc.numel() // (p.n_pixels * p.n_bins)
        if stream is None:
            stream = torch.cuda.current_stream(adc.device).cuda_stream
        f, i = self._f, self._i
        self.gen._handle.stats_accumulate(
            adc.data_ptr(), n_events,
            adc_hist=self._ptr(i, self._slices_i['adc_hist']),
            bin_sum=self._ptr(f, self._slices_f['bin_sum']), bin_sumsq=self._ptr(f, self._slices_f['bin_sumsq']),
            pixel_sum=self._ptr(f, self._slices_f['pixel_sum']), pixel_sumsq=self._ptr(f, self._slices_f['pixel_sumsq']),
            charge_hist=self._ptr(i, self._slices_i['charge_hist']) if self.with_charge else None,
            charge_lo=self.charge_lo, n_charge_bins=self.n_charge_bins,
            moments=self._ptr(f, self._slices_f['moments']), stream=stream)
        i[self._slices_i['n_events']] += n_events

    def generate(self, n_events, first_event=0, out=None, stream=None):
        """Generate ``n_events`` events with the statistics taken inside the generation kernel
        (``dct_generate_stats``): the cube lands in ``out`` as usual but is never read back."""
        torch = self._torch
        g, p = self.gen, self.gen.params
        dev = torch.device('cuda', g.device)
        n_events = int(n_events)
        if out is None:
            out = torch.empty((n_events, p.n_pixels, p.n_bins), dtype=torch.int16, device=dev)
        elif (out.dtype != torch.int16 or not out.is_contiguous() or out.device != dev
              or out.numel() != n_events * p.n_pixels * p.n_bins):
            raise ValueError('out must be a contiguous int16 CUDA tensor of n*P*B elements')
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        f, i = self._f, self._i
        if n_events:
            g._handle.generate_stats(
                g.seed, int(first_event), n_events, out.data_ptr(), None, None,
                adc_hist=self._ptr(i, self._slices_i['adc_hist']),
                bin_sum=self._ptr(f, self._slices_f['bin_sum']), bin_sumsq=self._ptr(f, self._slices_f['bin_sumsq']),
                pixel_sum=self._ptr(f, self._slices_f['pixel_sum']), pixel_sumsq=self._ptr(f, self._slices_f['pixel_sumsq']),
                charge_hist=self._ptr(i, self._slices_i['charge_hist']) if self.with_charge else None,
                charge_lo=self.charge_lo, n_charge_bins=self.n_charge_bins,
                moments=self._ptr(f, self._slices_f['moments']), stream=stream)
            i[self._slices_i['n_events']] += n_events
        return out

    def all_reduce(self, group=None):
        """Sum over ranks (the only collective of a sharded run; a few tens of KB)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self._f, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(self._i, op=dist.ReduceOp.SUM, group=group)

    def result(self):
        f, i = self._f.cpu().numpy(), self._i.cpu().numpy()
        out = {k: f[s].copy() for k, s in self._slices_f.items()}
        out.update({k: i[s].copy() for k, s in self._slices_i.items()})
        out['n_events'] = int(out['n_events'][0])
        return out

    def moments_nsb(self):
        """(mean, mean_error, std, std_error) over all samples, as ``compute_moments_nsb``
        defines them (reference nsb_baseline.py:21-32)."""
        r = self.result()
        return moments_from_power_sums(r['moments'], r['n_events'] * self.n_pixels * self.n_bins)


def moments_from_power_sums(power_sums, n):
    s1, s2, s3, s4 = (float(v) for v in power_sums)
    mean = s1 / n
    m2 = s2 / n - mean ** 2
    m4 = s4 / n - 4 * mean * s3 / n + 6 * mean ** 2 * s2 / n - 3 * mean ** 4
    std = np.sqrt(max(m2, 0.0))
    mean_error = std / np.sqrt(n)
    moment_2_error = np.sqrt(max(1. / n * (m4 - (n - 3) / (n - 1) * std ** 4), 0.0))
    std_error = 0.5 * moment_2_error / std if std > 0 else 0.0
    return mean, mean_error, std, std_error
