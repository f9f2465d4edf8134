# host-path ceiling and end-to-end bench at N = 1, 2, 4, 8 on one box (gpurun --gpus 8)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then python scripts/host_path.py; else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/host_path.py 2>/dev/null | grep '^{'; fi
done | tee gpurun_out/host_path.txt
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then python bench.py --steps 5 --warmup 3 --no-cpu-baseline; else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 2>/dev/null | grep '^{'; fi
done | tee gpurun_out/bench_scale_r02.txt | cut -c1-300
