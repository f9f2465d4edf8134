# For the record: the full C4 (10^6 events, 1 GPU) and per-level table; run with gpurun (1 GPU)
mkdir -p gpurun_out
python bench.py --steps 61 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_full_1e6.json 2>&1
tail -n 1 gpurun_out/bench_c4_full_1e6.json | python -c "
import sys, json
j = json.loads(sys.stdin.read())
n = j['steps'] * j['run_details']['events_per_step_per_gpu']
print('C4 full: %d events in %.2f s (%.3e events/s, %.3e samples/s), K1 %.2f ms/step' % (n, j['steps'] * j['ms_per_step'] / 1e3, j['events_per_s'], j['value'], j['roofline']['kernel_ms']))
"
python tests/tools/acdc_levels.py | tee gpurun_out/levels_table.md
