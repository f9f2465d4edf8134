# builds the library with extra nvcc flags ($1) on the GPU box and times the C4 / dark kernels
cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="$1" > /dev/null 2>&1 && cd ../..
echo "variant: $1"
for args in "--workload c4" "--workload dark"; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e $args 2>&1 | tail -1 | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('  $args', '| kernel_ms %.2f' % j['roofline']['kernel_ms'], '| ms/step %.2f' % j['ms_per_step'], j['config']['kernel'], 'mean %.4f' % j['validation']['mean_adc'])
"
done
