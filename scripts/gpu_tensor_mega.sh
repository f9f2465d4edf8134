# persistent (multi-group) tensor kernel against the CTA kernel: parity on 64 events + timing on 16384; run with gpurun
mkdir -p gpurun_out
for M in ${MEGA_LIST:-0 1 4}; do
  echo "=== DCT_TENSOR_MEGA=$M"
  DCT_TENSOR_MEGA=$M TC_EVENTS=64 TC_TIME_EVENTS=16384 timeout 300 python scripts/tensor_check.py 2>&1 | grep -E "analog diff|adc:|tensor:|rror|rap" | head -6
done
