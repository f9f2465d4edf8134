"""``NTraceGenerator``: the reference's generator API over the B200 kernels.

Mirrors the constructor / iterator / attribute surface of the reference class
(``digicamtoy/tracegenerator/__init__.py:38-226``) so that ``produce_data.py`` and the example
scripts run unchanged, while the per-event Monte Carlo (:228-381) runs in
``libdigicamtoy_sm100.so``.  Events are produced in device batches; the iterator hands them out
one at a time from pinned host memory.

Deliberate differences from the reference:
  * ``n_events=n`` yields exactly n events (the reference yields n+1, against its own test).
  * ADC counts saturate to ``saturate=(0, 4095)``; the reference wraps negatives and never clips.
  * ``voltage_drop`` scales ``n_photon`` per pixel (the reference broadcasts (P,1)x(P,) to (P,P)).
  * ``poisson=False`` uses ``round(n_photon)`` primaries (the reference raises on floats and
    lets the configured value grow from event to event).
  * the pe lists ``nsb_time / nsb_photon / mask`` are not stored by the generation kernels; reading
    them re-derives the current event's lists from the counter-based RNG (``trace_pe``).
"""
from __future__ import annotations

import os

import numpy as np

from . import native
from .params import ARTIFICIAL_BACKWARD_TIME, build_params, evaluate_template, resolve_pulse_shape_file

_KERNELS = {'auto': native.KERNEL_AUTO, 'scatter': native.KERNEL_SCATTER, 'sweep': native.KERNEL_SWEEP,
            'grid': native.KERNEL_GRID, 'tensor': native.KERNEL_TENSOR}


def _torch():
    import torch
    return torch


def _resolve_device(device):
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError('digicamtoy_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    if device is None:
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    dev = torch.device(device)
    if dev.type != 'cuda':
        raise ValueError('device must be a CUDA device, got {!r}'.format(device))
    return dev.index if dev.index is not None else torch.cuda.current_device()


class PinnedPool:
    """Page-locked host buffers, reused: pinning costs ~0.3 ms per MB, more than generating the data
    it will hold (a 238 MB dark.yml level: ~80 ms to pin, 20 ms to generate and copy)."""

    def __init__(self, max_bytes=16 << 30):
        import threading
        self._free, self._lock = {}, threading.Lock()
        self.max_bytes, self._held = int(max_bytes), 0       # idle buffers kept; more than that is handed back to the OS

    def acquire(self, shape, dtype):
        torch = _torch()
        n = 1
        for d in shape:
            n *= int(d)
        with self._lock:
            stack = self._free.get((n, dtype))
            buf = stack.pop() if stack else None
            if buf is not None:
                self._held -= buf.numel() * buf.element_size()
        if buf is None:
            buf = torch.empty(n, dtype=dtype).pin_memory()
        return buf.view(*shape)

    def release(self, *tensors):
        with self._lock:
            for t in tensors:
                if t is None:
                    continue
                nbytes = t.numel() * t.element_size()
                if self._held + nbytes <= self.max_bytes:
                    self._free.setdefault((t.numel(), t.dtype), []).append(t.reshape(-1))
                    self._held += nbytes

    def clear(self):
        with self._lock:
            self._free.clear()
            self._held = 0


PINNED = PinnedPool()


class NTraceGenerator:

    def __init__(self, time_start=0, time_end=200, time_sampling=4, n_pixels=1296, nsb_rate=0.6,
                 crosstalk=0.08, gain_nsb=True, n_photon=0, poisson=True, sigma_e=0.8, sigma_1=0.8,
                 gain=5.8, baseline=200, time_signal=20, jitter=0,
                 pulse_shape_file='/utils/pulse_SST-1M_pixel_0.dat', sub_binning=0, n_events=None,
                 voltage_drop=False,
                 # --- additions of this implementation (all optional)
                 seed=None, device=None, batch_events=None, saturate=(0, 4095), verbose=False,
                 kernel='auto', first_event=0, prefetch=True,
                 **kwargs):  # unknown keys are ignored: produce_data passes its whole option set (:113)
        p = build_params(time_start=time_start, time_end=time_end, time_sampling=time_sampling,
                         n_pixels=n_pixels, nsb_rate=nsb_rate, crosstalk=crosstalk, gain_nsb=gain_nsb,
                         n_photon=n_photon, poisson=poisson, sigma_e=sigma_e, sigma_1=sigma_1,
                         gain=gain, baseline=baseline, time_signal=time_signal, jitter=jitter,
                         pulse_shape_file=pulse_shape_file, sub_binning=sub_binning,
                         voltage_drop=voltage_drop)
        self.params = p
        # reference attribute names (row a11 of the survey)
        self.artificial_backward_time = ARTIFICIAL_BACKWARD_TIME
        self.n_samples = p.n_samples
        self.time_start, self.time_end, self.time_sampling = p.time_start, p.time_end, p.time_sampling
        self.n_pixels = p.n_pixels
        self.n_photon, self.time_signal, self.jitter = p.n_photon, p.time_signal, p.jitter
        self.poisson = p.poisson
        self.filename_pulse_shape = resolve_pulse_shape_file(pulse_shape_file)
        self.gain, self.sigma_e, self.sigma_1, self.baseline = p.gain, p.sigma_e, p.sigma_1, p.baseline
        self.gain_nsb, self.voltage_drop = gain_nsb, voltage_drop
        self.nsb_rate, self.crosstalk = p.nsb_rate, p.crosstalk
        self.tau, self.true_baseline = p.tau, p.true_baseline
        self.sub_binning = sub_binning
        self.n_events = n_events
        self.count = -1
        if verbose:   # the three constructor print-outs of the reference (:130-135, :159-164)
            print('Set NSB rate : {} [GHz] \tTrue NSB rate : {} [GHz]'.format(
                p.set_nsb_rate.mean(), p.nsb_rate.mean()))
            print('Set XT : {} [p.e.] \tTrue XT : {} [p.e.]'.format(
                p.set_crosstalk.mean(), p.crosstalk.mean()))
            print('Set Gain : {} [LSB/p.e.] \tTrue Gain : {} [LSB/p.e.]'.format(
                p.set_gain.mean(), p.gain.mean()))
            print('Set Baseline : {} [LSB]\tTrue Baseline : {} [LSB]\tTrue Baseline Shift : {} [LSB]'
                  .format(p.baseline.mean(), p.true_baseline.mean(),
                          p.true_baseline.mean() - p.baseline.mean()))

        if saturate is None:
            saturate = (-32768, 32767)
        self.saturate = (int(saturate[0]), int(saturate[1]))
        if kernel not in _KERNELS:
            raise ValueError('kernel must be one of {}'.format(sorted(_KERNELS)))
        self.seed = int.from_bytes(os.urandom(8), 'little') if seed is None else int(seed) & (2 ** 64 - 1)
        self.first_event = int(first_event)
        self.device = _resolve_device(device)
        self._handle = native.Handle(p, device=self.device, adc_min=self.saturate[0],
                                     adc_max=self.saturate[1], kernel=_KERNELS[kernel])
        self.kernel = self._handle.kernel
        if batch_events is None:
            # Iterator batches: ~128 MiB of ADC for long runs -- generation and copy overlap inside a
            # batch, 3.5e5 instead of 2.1e5 C4 events/s (scripts/iterator_batch.py) -- but ~32 MiB for
            # runs below 20000 events: pinning memory costs ~0.3 ms per MiB, as much as simulating
            # 10^4 events.
            long_run = n_events is None or int(n_events) >= 20000
            batch_events = max(1, min(4096, ((128 if long_run else 32) << 20) // (p.n_pixels * p.n_bins * 2)))
            # the first batches of a long run are small (the first event arrives after ~32 MiB, not 128)
            self._first_batch = max(1, min(batch_events, (32 << 20) // (p.n_pixels * p.n_bins * 2)))
        else:
            self._first_batch = int(batch_events)
        self.batch_events = int(batch_events)
        self._host = None          # (adc, time, charge) views of the current batch
        self._bufs = [None, None]  # the two pinned batch slots of the iterator
        self._prefetch = None      # (first event, slot, future) of the batch being generated ahead
        self._pool = None
        self.prefetch = bool(prefetch)
        self._batch_first = 0      # event index (relative) of slot 0 of the current batch
        self._batch_len = 0
        self.reset()

    # ------------------------------------------------------------------ reference surface
    def __str__(self):
        return ''.join('{}{}'.format(k, v) for k, v in sorted(self.__dict__.items())
                       if not k.startswith('_'))

    def __iter__(self):
        return self

    def __next__(self):
        self.next()
        return self

    @property
    def pulse_template(self):
        p = self.params
        return lambda t: evaluate_template(p.knot_x, p.coef, t)

    def get_pulse_shape(self, time, t_0=0, baseline=0, amplitude=1):
        return amplitude * self.pulse_template(np.asarray(time) - t_0) + baseline        # :186-188

    def reset(self):
        """State before the first event (:190-199): zeros, all simulated bins."""
        p = self.params
        self.cherenkov_time = np.zeros(p.n_photon.shape)
        self.cherenkov_photon = np.zeros(p.n_photon.shape, dtype=int)
        self.sampling_bins = np.arange(p.time_start, p.time_end, p.time_sampling)
        self.adc_count = np.zeros((p.n_pixels, self.sampling_bins.shape[0]))

    def next(self):
        if self.n_events is not None and self.count + 1 >= self.n_events:
            raise StopIteration
        self.count += 1
        slot = self.count - self._batch_first
        if self._host is None or slot >= self._batch_len or slot < 0:
            self._fill_batch(self.count)
            slot = 0
        adc, t, q = self._host
        self.adc_count = adc[slot]
        self.cherenkov_time = t[slot]
        self.cherenkov_photon = q[slot]
        self.sampling_bins = self.params.sampling_bins

    def __getattr__(self, name):
        # nsb_time / nsb_photon / mask (:248-257, :345-348) are never stored by the generation
        # kernels; on access they are re-derived for the current event from the same counters.
        if name in ('nsb_time', 'nsb_photon', 'mask') and 'count' in self.__dict__:
            if self.count < 0:
                P = self.params.n_pixels
                return {'nsb_time': np.zeros((P, 1)), 'nsb_photon': np.zeros((P, 1)),
                        'mask': np.zeros((P, 1))}[name]
            pe = self.trace_pe(1, first_event=self.first_event + self.count, noise=False)
            time, amp, n = pe['time'][0], pe['amplitude'][0], pe['count'][0]
            mask = (np.arange(time.shape[1])[None, :] < n[:, None]).astype(int)
            return {'nsb_time': time, 'nsb_photon': amp, 'mask': mask}[name]
        raise AttributeError(name)

    def trace_pe(self, n_events, first_event=None, noise=True):
        """NSB photo-electron clusters (and electronic noise) behind the traces of the given events,
        re-derived from the counter-based RNG: ``count (n,P)``, ``time (n,P,M) f64``,
        ``amplitude (n,P,M) f32`` (zero-padded), ``noise (n,P,B) f32``."""
        torch = _torch()
        p = self.params
        n_events = int(n_events)
        if first_event is None:
            first_event = self.first_event
        dev = torch.device('cuda', self.device)
        stream = torch.cuda.current_stream(dev).cuda_stream
        count = torch.zeros((n_events, p.n_pixels), dtype=torch.int32, device=dev)
        self._handle.trace_pe(self.seed, int(first_event), n_events, 0, count.data_ptr(), stream=stream)
        m = max(1, int(count.max().item())) if n_events else 1
        time = torch.zeros((n_events, p.n_pixels, m), dtype=torch.float64, device=dev)
        amp = torch.zeros((n_events, p.n_pixels, m), dtype=torch.float32, device=dev)
        nz = torch.zeros((n_events, p.n_pixels, p.n_bins), dtype=torch.float32, device=dev) if noise else None
        self._handle.trace_pe(self.seed, int(first_event), n_events, m, count.data_ptr(), time.data_ptr(),
                              amp.data_ptr(), nz.data_ptr() if noise else None, stream)
        torch.cuda.synchronize(dev)
        out = dict(count=count.cpu().numpy(), time=time.cpu().numpy(), amplitude=amp.cpu().numpy())
        if noise:
            out['noise'] = nz.cpu().numpy()
        return out

    # ------------------------------------------------------------------ bulk API (device / host)
    def generate(self, n_events, first_event=None, out=None, truth=False, stream=None):
        """ADC cube of ``n_events`` events as a CUDA int16 tensor ``[n, P, B]`` (asynchronous).

        ``first_event`` is the global event index (default: continue after the last bulk call);
        the same (seed, event, pixel) always gives the same trace, whatever the batching or the
        number of GPUs.  With ``truth=True`` also returns (time, charge) float32 ``[n, P, K]``."""
        torch = _torch()
        p = self.params
        n_events = int(n_events)
        if first_event is None:
            first_event = self.first_event + getattr(self, '_bulk_cursor', 0)
            self._bulk_cursor = getattr(self, '_bulk_cursor', 0) + n_events
        dev = torch.device('cuda', self.device)
        if out is None:
            out = torch.empty((n_events, p.n_pixels, p.n_bins), dtype=torch.int16, device=dev)
        elif (out.dtype != torch.int16 or not out.is_contiguous() or out.device != dev
              or out.numel() != n_events * p.n_pixels * p.n_bins):
            raise ValueError('out must be a contiguous int16 CUDA tensor of n*P*B elements')
        t = q = None
        if truth:
            t = torch.empty((n_events, p.n_pixels, p.n_pulses), dtype=torch.float32, device=dev)
            q = torch.empty_like(t)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        if n_events:
            self._handle.generate(self.seed, int(first_event), n_events, out.data_ptr(),
                                  t.data_ptr() if truth else None, q.data_ptr() if truth else None,
                                  stream)
        return (out, t, q) if truth else out

    def generate_analog(self, n_events, first_event=0):
        """Test hook: ``(adc int16 [n,P,B], analog float32 [n,P,B])`` CUDA tensors, ``analog`` being the
        p.e.-weighted template sums each sample is digitised from (before gain, noise, baseline)."""
        torch = _torch()
        p = self.params
        dev = torch.device('cuda', self.device)
        adc = torch.empty((int(n_events), p.n_pixels, p.n_bins), dtype=torch.int16, device=dev)
        analog = torch.empty((int(n_events), p.n_pixels, p.n_bins), dtype=torch.float32, device=dev)
        if n_events:
            self._handle.generate_analog(self.seed, int(first_event), int(n_events), adc.data_ptr(),
                                         analog.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        return adc, analog

    def generate_host(self, n_events, first_event=None, out=None, truth=True, pool=None):
        """Same, into pinned host memory through the library's chunked D2H pipeline; returns
        NumPy views ``(adc uint16/int16 [n,P,B], time f32 [n,P,K], charge f32 [n,P,K])``.
        ``pool`` (a ``PinnedPool``): take the buffers from it instead of pinning fresh memory; the
        tensors are then in ``self.last_host_tensors`` for the caller to ``release`` when done."""
        torch = _torch()
        p = self.params
        n_events = int(n_events)
        if first_event is None:
            first_event = self.first_event
        if out is None:
            adc = (pool.acquire((n_events, p.n_pixels, p.n_bins), torch.int16) if pool is not None else
                   torch.empty((n_events, p.n_pixels, p.n_bins), dtype=torch.int16).pin_memory())
        else:
            adc = out
            if (not isinstance(adc, torch.Tensor) or adc.dtype != torch.int16 or adc.is_cuda
                    or not adc.is_contiguous() or adc.numel() != n_events * p.n_pixels * p.n_bins):
                raise ValueError('out must be a contiguous int16 CPU tensor of n*P*B elements')
        t = q = None
        if truth and pool is not None:
            t = pool.acquire((n_events, p.n_pixels, p.n_pulses), torch.float32)
            q = pool.acquire((n_events, p.n_pixels, p.n_pulses), torch.float32)
        elif truth:
            t = torch.empty((n_events, p.n_pixels, p.n_pulses), dtype=torch.float32).pin_memory()
            q = torch.empty((n_events, p.n_pixels, p.n_pulses), dtype=torch.float32).pin_memory()
        if self._prefetch is not None:                   # the host pipeline of a handle serves one call at a time
            self._prefetch[2].result()
        if n_events:
            self._handle.generate_host(self.seed, int(first_event), n_events, adc.data_ptr(),
                                       t.data_ptr() if truth else None,
                                       q.data_ptr() if truth else None)
        self._pinned_keepalive = (adc, t, q)
        self.last_host_tensors = (adc if out is None else None, t, q)
        a = adc.numpy()
        if self.saturate[0] >= 0:
            a = a.view(np.uint16)
        return (a, t.numpy(), q.numpy()) if truth else a

    # ------------------------------------------------------------------ internals
    def _batch_buffers(self, which, n_needed):
        """Pinned (adc, time, charge) tensors of iterator batch slot `which` (0 / 1) for n_needed events."""
        torch = _torch()
        if self._bufs[which] is None or self._bufs[which][0].shape[0] < n_needed:
            p, n = self.params, int(n_needed)
            if self._bufs[which] is not None:                    # growing: straight to the full batch size, once per slot
                n = max(n, self.batch_events if self.n_events is None else min(self.batch_events, int(self.n_events)))
            self._bufs[which] = (torch.empty((n, p.n_pixels, p.n_bins), dtype=torch.int16).pin_memory(),
                                 torch.empty((n, p.n_pixels, p.n_pulses), dtype=torch.float32).pin_memory(),
                                 torch.empty((n, p.n_pixels, p.n_pulses), dtype=torch.float32).pin_memory())
        return self._bufs[which]

    def _generate_batch(self, which, first_relative, n):
        """Batch [first_relative, +n) into slot `which`; returns NumPy views like generate_host."""
        adc, t, q = self._batch_buffers(which, n)
        self._handle.generate_host(self.seed, self.first_event + first_relative, n, adc.data_ptr(),
                                   t.data_ptr(), q.data_ptr())
        a = adc.numpy()[:n]
        if self.saturate[0] >= 0:
            a = a.view(np.uint16)
        return a, t.numpy()[:n], q.numpy()[:n]

    def _batch_size_at(self, first_relative):
        n = min(self.batch_events, max(self._first_batch, int(first_relative)))     # 1, 1, 2, 4 ... first batches
        if self.n_events is not None:
            n = max(0, min(n, self.n_events - first_relative))
        return n

    def _fill_batch(self, first_relative):
        """The iterator's batch that starts at event `first_relative`: two pinned slots, reused; while
        the caller walks through one, the next batch is generated and copied into the other on a
        worker thread (the C call releases the GIL).  The per-event arrays the iterator hands out are
        views into a slot and stay valid until that slot is refilled, one full batch later."""
        n = self._batch_size_at(first_relative)
        pending = self._prefetch
        self._prefetch = None
        if pending is not None and pending[0] == first_relative:
            which, host = pending[1], pending[2].result()
        else:
            if pending is not None:
                pending[2].result()                      # never two calls on one handle at a time
            which = 0
            host = self._generate_batch(which, first_relative, n)
        self._host = host
        self._batch_first = first_relative
        self._batch_len = n
        nxt = first_relative + n
        n_next = self._batch_size_at(nxt)
        if n_next > 0 and self.prefetch:
            if self._pool is None:
                from concurrent.futures import ThreadPoolExecutor
                self._pool = ThreadPoolExecutor(max_workers=1)
            self._prefetch = (nxt, 1 - which, self._pool.submit(self._generate_batch, 1 - which, nxt, n_next))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if getattr(self, '_prefetch', None) is not None:
            try:
                self._prefetch[2].result()
            except Exception:
                pass
            self._prefetch = None
        if getattr(self, '_pool', None) is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        if getattr(self, '_handle', None) is not None:
            self._handle.close()
