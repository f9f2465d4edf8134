# on-grid kernel (K1c) under build variants: bit-identity with the scatter kernel + timing
for V in "$@"; do
  (cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="$V" > /dev/null 2>&1)
  echo "=== variant: $V"
  python -m pytest tests/test_generate_gpu.py -q -k "grid_and_scatter" 2>&1 | tail -1
  python scripts/time_kernels.py 2>&1 | grep grid
done
