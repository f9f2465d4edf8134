mkdir -p gpurun_out
python -m pytest tests/test_produce_and_shards_gpu.py -x -q 2>&1 | grep -v "^INFO" | tail -30
python scripts/run_k3_once.py dark > gpurun_out/plain_k3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_stats -s 1 -c 1 -o gpurun_out/prof_k3_dark -f python scripts/run_k3_once.py dark > gpurun_out/ncu_k3.log 2>&1
tail -n 2 gpurun_out/ncu_k3.log
for args in "--workload c4"; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e $args 2>&1 | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('$args', '| value %.3e' % j['value'], '| ms/step %.2f' % j['ms_per_step'], '| kernel_ms %.2f' % j['roofline']['kernel_ms'], '| GB/s %.1f' % j['roofline']['achieved'], '| ev/s %.3e' % j['events_per_s'], j['config']['kernel'], j['validation'])
"
done
