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
        self._torch = torch
        self.gen = generator
        p = generator.params
        lo, hi = generator.saturate
        dev = torch.device('cuda', generator.device)
        self.charge_lo, self.n_charge_bins = int(charge_lo), int(n_charge_bins)
        # one flat float64 + one flat int64 buffer: a sharded run all-reduces exactly two tensors
        self._f = torch.zeros(2 * p.n_bins + 2 * p.n_pixels + 4, dtype=torch.float64, device=dev)
        self._i = torch.zeros((hi - lo + 1) + self.n_charge_bins + 1, dtype=torch.int64, device=dev)
        self._slices_f = {}
        o = 0
        for name, n in (('bin_sum', p.n_bins), ('bin_sumsq', p.n_bins), ('pixel_sum', p.n_pixels),
                        ('pixel_sumsq', p.n_pixels), ('moments', 4)):
            self._slices_f[name] = slice(o, o + n)
            o += n
        self._slices_i = {'adc_hist': slice(0, hi - lo + 1),
                          'charge_hist': slice(hi - lo + 1, hi - lo + 1 + self.n_charge_bins),
                          'n_events': slice(hi - lo + 1 + self.n_charge_bins, hi - lo + 2 + self.n_charge_bins)}
        self.with_charge = p.n_bins <= 128

    def _ptr(self, buf, sl):
        return buf.data_ptr() + sl.start * buf.element_size()

    def accumulate(self, adc, stream=None):
        torch = self._torch
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
        return moments_from_power_sums(self.result()['moments'],
                                       self.result()['n_events'] * self.gen.params.n_pixels * self.gen.params.n_bins)


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
