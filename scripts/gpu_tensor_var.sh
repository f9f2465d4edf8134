# tensor kernel under build variants ($@ = one EXTRA string per variant): parity on 64 events + timing on 16384
# (MEGA = value of DCT_TENSOR_MEGA for the run: 0 CTA kernel, 1 persistent, 4 persistent with four groups)
mkdir -p gpurun_out
for V in "$@"; do
  (cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="$V" > /dev/null 2>&1)
  echo "=== variant: $V  (DCT_TENSOR_MEGA=${MEGA:-1})"
  DCT_TENSOR_MEGA=${MEGA:-1} TC_EVENTS=64 TC_TIME_EVENTS=16384 timeout 300 python scripts/tensor_check.py 2>&1 | grep -E "analog diff|adc:|tensor:|rror:" | head -4
done
