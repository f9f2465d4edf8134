"""What the box's host path sustains: N ranks (torchrun), each doing nothing but cudaMemcpyAsync D2H of
one C4 step (2.12 GB) from its GPU into its own pinned buffer, with the NUMA binding bench.py uses.
The aggregate GB/s is the ceiling of bench.py's end-to-end number at that N.
Also times the same copy in 64 MiB pieces (what dct_generate_host issues)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import bind_to_gpu_numa_node

world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
cpus = bind_to_gpu_numa_node(local) if world > 1 else 'all'
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
n = 16384 * 1296 * 50
src = torch.zeros(n, dtype=torch.int16, device=dev)
dst = torch.empty(n, dtype=torch.int16).pin_memory()
stream = torch.cuda.Stream(dev)

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)

def timed(fn, reps=5):
    fn(); barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(dev)
    ms = (time.perf_counter() - t0) * 1e3 / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms

def whole():
    with torch.cuda.stream(stream):
        dst.copy_(src, non_blocking=True)
    stream.synchronize()

chunk = (64 << 20) // 2
def pieces():
    with torch.cuda.stream(stream):
        for o in range(0, n, chunk):
            dst[o:o + chunk].copy_(src[o:o + chunk], non_blocking=True)
    stream.synchronize()

ms_w, ms_p = timed(whole), timed(pieces)
if rank == 0:
    gb = n * 2 / 1e9
    print(json.dumps(dict(n_gpus=world, bytes_per_rank=n * 2, ms_whole=ms_w, ms_64MiB_pieces=ms_p,
                          per_rank_GBps=gb / (ms_w * 1e-3), aggregate_GBps=world * gb / (ms_w * 1e-3),
                          aggregate_GBps_pieces=world * gb / (ms_p * 1e-3),
                          cpus_rank0=(cpus if isinstance(cpus, str) else '%d cpus: %s..%s' % (len(cpus), cpus[0], cpus[-1])))))
if world > 1:
    dist.destroy_process_group()
