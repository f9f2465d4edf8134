# which accumulator column offsets does tcgen05.mma accept as D?  (each variant in its own process)
mkdir -p gpurun_out
for A in 32 16 8 4 2; do
  (cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="-DDCT_TC_DBG_COL_ALIGN=$A" > /dev/null 2>&1)
  echo "=== align $A"
  TC_EVENTS=4 TC_TIME=0 timeout 120 python scripts/tensor_check.py 2>&1 | grep -E "analog diff|adc:|Error|error" | head -4
done 2>&1 | tee gpurun_out/tensor_align.txt
