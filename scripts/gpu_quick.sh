# quick GPU check: tests + short bench lines (kernel-only numbers) + a few pulse-heavy levels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for args in "--workload c4" "--workload dark"; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e $args 2>&1 | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('$args', '| value %.3e' % j['value'], '| ms/step %.2f' % j['ms_per_step'], '| kernel_ms %.2f' % j['roofline']['kernel_ms'], '| GB/s %.1f' % j['roofline']['achieved'], '| ev/s %.3e' % j['events_per_s'], j['config']['kernel'])
"
done
python - <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from digicamtoy_b200 import NTraceGenerator, workloads
E = 8192
for rate, npe in ((0, 0.), (0, 100), (0.003, 100), (0.04, 100), (1, 100)):
    g = NTraceGenerator(seed=1, **workloads.c2_ac_dc(rate, npe))
    out = torch.empty((E, 1296, 50), dtype=torch.int16, device='cuda')
    for _ in range(2): g.generate(E, first_event=0, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3): g.generate(E, first_event=(i + 1) * E, out=out)
    e1.record(); torch.cuda.synchronize()
    print('ac_dc nsb %g npe %g: %.3e events/s' % (rate, npe, 3 * E / (e0.elapsed_time(e1) * 1e-3)))
PY
