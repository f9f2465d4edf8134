# weak-scaling check on one box: N = 1, 2, 4, 8 (torchrun for N > 1), short runs
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scale_n1.json 2> gpurun_out/bench_scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_scale_n$n.json 2> gpurun_out/bench_scale_n$n.err
done
for n in 1 2 4 8; do tail -n 1 gpurun_out/bench_scale_n$n.json | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('N=%d value %.4e ms/step %.2f e2e %.4e reduced %d' % (j['n_gpus'], j['value'], j['ms_per_step'], j['e2e']['value'], j['validation']['n_events_reduced']))
"; done
