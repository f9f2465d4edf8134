mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "stats or chunking or bit_identical or saturation" 2>&1 | tail -5
python scripts/time_stats.py
for args in "--workload c4" "--workload dark"; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e $args 2>&1 | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('$args', '| value %.3e' % j['value'], '| ms/step %.2f' % j['ms_per_step'], '| kernel_ms %.2f' % j['roofline']['kernel_ms'], '| GB/s %.1f' % j['roofline']['achieved'], '| ev/s %.3e' % j['events_per_s'], j['config']['kernel'], j['validation'])
"
done
