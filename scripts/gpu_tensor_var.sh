# tensor kernel under build variants ($@ = one EXTRA string per variant): parity on 64 events + timing on 16384
mkdir -p gpurun_out
for V in "$@"; do
  (cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="$V" > /dev/null 2>&1)
  echo "=== variant: $V"
  TC_EVENTS=64 TC_TIME_EVENTS=16384 timeout 300 python scripts/tensor_check.py 2>&1 | grep -E "analog diff|adc:|tensor:|rror:" | head -4
done
