mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "stats or chunking or bit_identical or saturation" 2>&1 | tail -5
for args in "--workload c4" "--workload dark"; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e $args 2>&1 | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('$args', '| value %.3e' % j['value'], '| ms/step %.2f' % j['ms_per_step'], '| kernel_ms %.2f' % j['roofline']['kernel_ms'], '| GB/s %.1f' % j['roofline']['achieved'], '| ev/s %.3e' % j['events_per_s'], j['config']['kernel'], j['validation'])
"
done
python bench.py --steps 2 --warmup 1 --events-per-step 4096 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_generate -s 1 -c 1 -o gpurun_out/prof_c4_sweep_v2 -f python bench.py --steps 2 --warmup 1 --events-per-step 4096 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/ncu_full.log
