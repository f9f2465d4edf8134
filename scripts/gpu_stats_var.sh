# statistics kernel under build variants ($@ = EXTRA strings): step and K1 times of the c4 / acdc / grid / dark benches
mkdir -p gpurun_out
for V in "$@"; do
  (cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="$V" > /dev/null 2>&1)
  echo "=== variant: $V"
  for w in c4 acdc grid dark; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print(sys.argv[1], 'step %.2f ms  K1 %.2f ms  K3 %.2f ms' % (j['ms_per_step'], j['roofline']['kernel_ms'], j['ms_per_step'] - j['roofline']['kernel_ms']), j['validation']['mean_adc'])" $w; done
done
(cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make > /dev/null 2>&1)
