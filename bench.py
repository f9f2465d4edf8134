#!/usr/bin/env python
"""Benchmark of the trace-generation hot path (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (N=1): BASELINE.json config C4 -- full camera, 1296 px x 50 samples, 1 GHz NSB on every
pixel, one signal pulse per pixel (1..3162 pe, top ~10 % of pixels saturate 12 bits).  A "step" is
one pass of the hot path over one batch of `--events-per-step` events (default 16384 = 2.1 GB of
ADC, far larger than the 126 MB L2) followed by the on-device validation statistics of that batch.
At N>1 (torchrun, one rank per GPU) every rank generates its own contiguous range of the global
event index (weak scaling, no data-path collective) and the statistics are all-reduced once at
the end of the timed region (NCCL).

Printed JSON line (rank 0): metric "digitised ADC samples/s", `value` device-resident throughput,
`e2e` the same through the host-buffer C-ABI call (D2H inside the timed region), `roofline` of the
generation kernel against the measured HBM copy bandwidth, `cpu_baseline` the oracle on host cores.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'digitised ADC samples/s (1296 px x 50 samples, 1 GHz NSB + signal, 12-bit saturation)'
UNIT = 'samples/s'
WORKLOAD = 'C4 full camera: 1296 px, 0-200 ns / 4 ns (50 bins), NSB 1 GHz, xt 0.08, n_photon 1..3162 pe'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--events-per-step', type=int, default=16384)
    ap.add_argument('--kernel', default='auto', choices=['auto', 'scatter', 'sweep', 'tensor', 'grid'])
    ap.add_argument('--workload', default='c4', choices=['c4', 'dark', 'acdc', 'grid'])
    ap.add_argument('--fused-stats', action='store_true',
                    help='statistics taken inside the generation kernel (dct_generate_stats) instead of by the stand-alone '
                         'kernel.  Measured slower (dark.yml 8.7 against 7.2 ms per step, C4 62.3 against 61.9): the '
                         'statistics are instruction-bound, not HBM-bound, so the saved read of the cube buys nothing')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--cpu-events', type=int, default=3, help='timed oracle events per host worker')
    return ap.parse_args()


def workload_kwargs(name):
    from digicamtoy_b200 import workloads
    if name == 'dark':
        return workloads.c1_dark(n_photon=0.), 'C1 dark.yml: 1296 px, 92 bins, NSB 3 MHz'
    if name == 'acdc':
        return workloads.c2_ac_dc(0.125, 10.), 'C2 ac_dc.yml level: 1296 px, 50 bins, NSB 0.125 GHz, n_photon 10'
    if name == 'grid':
        kw = workloads.c4_full_camera()
        kw['sub_binning'] = 1
        return kw, 'C4 geometry with on-grid NSB (sub_binning=1): 1296 px, 50 bins, NSB 1 GHz, n_photon 1..3162 pe'
    return workloads.c4_full_camera(), WORKLOAD


# --------------------------------------------------------------------------------- CPU oracle legs
def _oracle_worker(args):
    """One host core: build the oracle generator, one warm-up event, then timed events."""
    kwargs, seed, n_warm, n_timed = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    import warnings
    warnings.simplefilter('ignore')
    from oracle import reference_port as orc
    gen = orc.OracleGenerator(seed=seed, **kwargs)
    for _ in range(n_warm):
        next(gen)
    t = time.perf_counter()
    for _ in range(n_timed):
        next(gen)
    return time.perf_counter() - t


def host_cores(cap=32):
    """Cores this process may use; capped because every oracle worker holds the reference's dense
    (pixel, pe, bin) arrays (~0.6 GB at 1 GHz NSB)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, cap))


def time_oracle(kwargs, n_warm, n_timed, cores):
    """All host cores, one single-threaded oracle per core (the reference is single-threaded
    NumPy/SciPy).  Returns (samples/s summed over workers, wall seconds)."""
    ctx = mp.get_context('spawn')
    jobs = [(kwargs, 1000 + i, n_warm, n_timed) for i in range(cores)]
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        secs = pool.map(_oracle_worker, jobs)
    wall = time.perf_counter() - t0
    samples_per_event = kwargs['n_pixels'] * int((kwargs['time_end'] - kwargs['time_start']) // kwargs['time_sampling'])
    rate = sum(n_timed * samples_per_event / s for s in secs)
    return rate, wall, secs


def subset_pixels(kwargs, n_pixels):
    """The same workload on the first n_pixels pixels (per-pixel lists are cut, scalars kept)."""
    kw = dict(kwargs)
    full = kw['n_pixels']
    kw['n_pixels'] = n_pixels
    for k, v in kwargs.items():
        if isinstance(v, (list, tuple)) and len(v) == full:
            kw[k] = list(v[:n_pixels])
    return kw


def run_reference_arm(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    kwargs, name = workload_kwargs(a.workload)
    cores = host_cores()
    # One oracle event per worker per "step".  A full-camera C4 event costs ~3.8 s on one core; if
    # steps + warm-up would run past ~150 s the sample is cut to a leading block of pixels
    # (cost per pixel is constant, so samples/s is unaffected).
    sec_per_event = {'c4': 3.8, 'grid': 3.8, 'acdc': 0.6}.get(a.workload, 0.12)
    est = sec_per_event * (a.steps + a.warmup)
    n_px = kwargs['n_pixels']
    kwargs_full_pixels = n_px
    n_bins_cfg = int((kwargs['time_end'] - kwargs['time_start']) // kwargs['time_sampling'])
    if est > 150.0:
        n_px = max(81, int(n_px * 150.0 / est))
        kwargs = subset_pixels(kwargs, n_px)
    rate, wall, secs = time_oracle(kwargs, a.warmup, a.steps, cores)
    per_step_ms = 1e3 * max(secs) / max(a.steps, 1)
    sample = '%d step(s) of 1 event x %d px per worker after %d warm-up, %d single-threaded workers, %.1f s wall' % (
        a.steps, n_px, a.warmup, cores, wall)
    line = dict(impl='reference', metric=METRIC, value=rate, unit=UNIT, n_gpus=a.gpus, steps=a.steps,
                warmup=a.warmup, ms_per_step=per_step_ms, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f64', data='synthetic',
                config=dict(workload=name, n_pixels=kwargs_full_pixels, n_bins=n_bins_cfg),
                run_details=dict(events_per_step=cores, pixels_per_event=n_px,
                                 timing='time.perf_counter around the oracle loop, per worker; value = sum over workers'),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=cores, kind='port', sample=sample),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line))


# --------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, mx, reasons, power = [], 0, set(), []
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for t, row in self.rows:
            if not (t0 <= t <= t1):
                continue
            f = [x.strip() for x in row.split(',')]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1])); power.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None,
                    power_w_max=max(power) if power else None, samples=len(sm), reasons=sorted(reasons))


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs next to its GPU before any pinned buffer is allocated, so that the
    D2H copies of the end-to-end path do not cross the socket interconnect (8 ranks on one box share
    host memory bandwidth; first-touch puts the pinned pages on the local node)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:      # match by PCI address: NVML and CUDA may enumerate in different orders
            import torch
            pr = torch.cuda.get_device_properties(index)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(
                '%08x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return sorted(os.sched_getaffinity(0))
    except Exception as e:      # no NVML, restricted cpuset, ...: keep the inherited affinity
        return 'unchanged (%s)' % type(e).__name__


# --------------------------------------------------------------------------------- the B200 arm
def run_b200_arm(a):
    import torch
    import torch.distributed as dist
    from digicamtoy_b200 import NTraceGenerator, workloads
    from digicamtoy_b200.stats import CubeStats

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the B200 arm has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    cpus = bind_to_gpu_numa_node(local) if world > 1 else 'all'
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    kwargs, name = workload_kwargs(a.workload)
    gen = NTraceGenerator(seed=workloads.C4_SEED, device=local, kernel=a.kernel, **kwargs)
    P, B = gen.params.n_pixels, gen.params.n_bins
    E = a.events_per_step
    stats = CubeStats(gen)
    ring = [torch.empty((E, P, B), dtype=torch.int16, device=dev) for _ in range(2)]
    stream = torch.cuda.current_stream(dev)

    def step(i, record=None):
        # weak scaling: global step i covers events [(i*world + rank)*E, +E)
        first = (i * world + rank) * E
        out = ring[i & 1]
        if record is not None:
            record[0].record(stream)
        if (not a.fused_stats):
            gen.generate(E, first_event=first, out=out)
        else:
            stats.generate(E, first_event=first, out=out)       # statistics fused into the generation kernel
        if record is not None:
            record[1].record(stream)
        if (not a.fused_stats):
            stats.accumulate(out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(a.warmup):
        step(i)
    barrier()
    stats._f.zero_(); stats._i.zero_()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    kernel_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(a.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    t_begin.record(stream)
    for i in range(a.steps):
        step(a.warmup + i, kernel_events[i])
    stats.all_reduce()
    t_end.record(stream)
    barrier()
    w1 = time.perf_counter()
    elapsed_ms = t_begin.elapsed_time(t_end)
    kernel_ms = [e0.elapsed_time(e1) for e0, e1 in kernel_events]
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    samples_per_step = E * P * B * world
    value = samples_per_step * a.steps / (elapsed_ms * 1e-3)

    # ---- end to end: host buffers through the C ABI (dct_generate_host), D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        host = torch.empty((E, P, B), dtype=torch.int16).pin_memory()
        n_e2e = max(1, min(a.steps, 5))
        gen.generate_host(E, first_event=rank * E, out=host, truth=False)      # warm-up (allocates the pipeline)
        barrier()
        h0 = time.perf_counter()
        for i in range(n_e2e):
            gen.generate_host(E, first_event=((a.warmup + i) * world + rank) * E, out=host, truth=False)
        torch.cuda.synchronize(dev)
        h_ms = (time.perf_counter() - h0) * 1e3
        if world > 1:
            t = torch.tensor([h_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            h_ms = float(t.item())
        e2e = dict(value=samples_per_step * n_e2e / (h_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=0,
                   d2h_bytes_per_step=E * P * B * 2, steps=n_e2e, ms_per_step=h_ms / n_e2e,
                   api='NTraceGenerator.generate_host -> dct_generate_host (pinned host buffer, chunked D2H overlapped with generation)',
                   host_cpus_rank0=(cpus if isinstance(cpus, str) else '%d cpus: %s..%s' % (len(cpus), cpus[0], cpus[-1])))
        # what a plain N-way cudaMemcpyAsync D2H of the same bytes sustains on this kind of box
        # (scripts/host_path.py, profiles/r02/host_path.txt): the ceiling of any end-to-end number
        cpath = os.path.join(ROOT, 'profiles', 'r02', 'host_path_ceiling.json')
        if os.path.exists(cpath):
            ceiling = json.load(open(cpath)).get(str(world))
            if ceiling:
                gbps = e2e['value'] * 2 / 1e9
                e2e.update(GBps=gbps, plain_copy_ceiling_GBps=ceiling, frac_of_ceiling=gbps / ceiling,
                           bound='host ingest: PCIe link at N = 1, the box\'s aggregate host path beyond.  The ceiling is what N plain D2H copies '
                                 'reached on one of the pool\'s 8-GPU boxes (profiles/r02/host_path.txt); it varies by box (a 2-GPU box '
                                 'gave 87 GB/s at N = 2), so the fraction can exceed 1')
        del host
    # ---- the reference's own way of consuming the generator: `for event in generator`
    # (reference produce_data.py:119-126), one event at a time out of prefetched pinned batches
    iterator = None
    if rank == 0 and not a.no_e2e:
        n_it = 8192 if B <= 64 else 4096
        it_gen = NTraceGenerator(seed=workloads.C4_SEED + 1, device=local, kernel=a.kernel, **kwargs)
        n_warm = 2 * it_gen.batch_events + 1           # both pinned batch slots and the host pipeline exist after this
        sink = 0
        for _ in range(n_warm):
            sink += int(next(it_gen).adc_count[0, 0])
        i0 = time.perf_counter()
        for _ in range(n_it):
            sink += int(next(it_gen).adc_count[0, 0])  # touch the event like a consumer would
        i_s = time.perf_counter() - i0
        iterator = dict(events_per_s=n_it / i_s, events=n_it, warmup_events=n_warm, batch_events=it_gen.batch_events,
                        api='for event in NTraceGenerator(...): event.adc_count (two pinned batch slots, next batch prefetched)')
        it_gen.close()
    if rank == 0:
        clocks.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K1)
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    k_ms = sum(kernel_ms) / len(kernel_ms)
    algo_bytes = E * P * B * 2
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    traffic, sm_bound, kernel_name = None, None, 'k_generate_' + gen.kernel
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(gen.kernel)
        if tj:
            if not a.fused_stats:          # (with fused statistics K1t runs as the CTA kernel)
                kernel_name = tj.get('kernel_name', kernel_name)
            per_sample = tj.get(a.workload, tj)['dram_bytes_per_sample'] if a.workload != 'c4' else tj['dram_bytes_per_sample']
            traffic = per_sample * E * P * B
            if a.workload == 'c4' and 'sm_issue_active_pct' in tj:
                # the resources that actually bind this kernel, from the committed ncu capture
                sm_bound = dict(issue_active_pct_of_peak=tj['sm_issue_active_pct'],
                                shared_wavefronts_pct_of_peak=tj['shared_wavefronts_pct_of_peak'],
                                active_threads_per_warp_inst=tj['active_threads_per_warp_inst'],
                                source=tj.get('ncu'))
    roofline = dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved / peak,
                    traffic=traffic, kernel=kernel_name, kernel_ms=k_ms, sm_bound=sm_bound,
                    algorithmic_bytes_per_launch=algo_bytes, peak_source=peak_src,
                    note='HBM-write roofline (2 B/sample). This workload is SM-issue-bound, not HBM-bound: '
                         '~245 photo-electron clusters per trace x 22 template taps; see DESIGN.md')

    stats_result = stats.result()
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup,
                ms_per_step=elapsed_ms / a.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic',
                config=dict(workload=name, n_pixels=P, n_bins=B),
                run_details=dict(events_per_step_per_gpu=E, kernel=gen.kernel, seed=workloads.C4_SEED,
                                 statistics='separate kernel' if (not a.fused_stats) else 'fused into the generation kernel',
                                 l2='each step writes %.2f GB per GPU into a 2-deep ring (>> 126 MB L2); no input re-read' % (algo_bytes / 1e9),
                                 sharding='contiguous global-event ranges per rank; one all-reduce of the statistics at the end'),
                events_per_s=value / (P * B), roofline=roofline, e2e=e2e, iterator=iterator,
                iterator_events_per_s=iterator['events_per_s'] if iterator else None,
                gpu_launches=(2 if (not a.fused_stats) else 1) * a.steps,
                gpu_launches_detail=('per step: 1 x %s + 1 x k_stats_tile' % kernel_name if (not a.fused_stats) else
                                     'per step: 1 x k_generate_%s (statistics fused)' % gen.kernel),
                clocks=clocks.summary(w0, w1),
                validation=dict(mean_adc=stats_result['moments'][0] / max(1, stats_result['n_events'] * P * B),
                                n_events_reduced=stats_result['n_events'],
                                frac_saturated=float(stats_result['adc_hist'][-1]) / max(1, stats_result['n_events'] * P * B)))

    if world == 1 and not a.no_cpu_baseline:
        cores = host_cores()
        rate, wall, secs = time_oracle(kwargs, 1, a.cpu_events, cores)
        line['cpu_baseline'] = dict(value=rate, unit=UNIT, cores=cores, kind='port',
                                    sample='%d event(s) x %d px per worker after 1 warm-up event, %d single-threaded workers, %.1f s wall'
                                           % (a.cpu_events, P, cores, wall),
                                    single_core_samples_per_s=rate / cores)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == 'reference':
        run_reference_arm(a)
        return
    if a.gpus > 1 and 'WORLD_SIZE' not in os.environ:
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(a.gpus),
               '--master-addr', '127.0.0.1', '--master-port', '29533'] + sys.argv
        raise SystemExit(subprocess.call(cmd))
    run_b200_arm(a)


if __name__ == '__main__':
    main()
