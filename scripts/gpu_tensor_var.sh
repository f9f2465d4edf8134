# tensor kernel under build variants ($@ = one EXTRA string per variant), 64 and 1024 events
mkdir -p gpurun_out
for V in "$@"; do
  (cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="$V" > /dev/null 2>&1)
  echo "=== variant: $V"
  for ev in 64 1024; do
    TC_EVENTS=$ev TC_TIME=0 CUDA_LAUNCH_BLOCKING=1 timeout 120 python scripts/tensor_check.py 2>&1 | grep -E "analog diff|adc:|rror:" | head -3
  done
done
