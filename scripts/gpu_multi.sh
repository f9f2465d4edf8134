# 2-GPU check of the torchrun path (NCCL all-reduce of the statistics), both arms
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -c 400
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -c 1200
