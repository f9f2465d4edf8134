# tensor-core kernel: parity against the sweep kernel + timing, each step in its own bounded process
mkdir -p gpurun_out
for ev in 4 16 64; do
  echo "=== parity, $ev events"
  ( time TC_EVENTS=$ev TC_TIME=0 CUDA_LAUNCH_BLOCKING=1 timeout 120 python scripts/tensor_check.py 2>&1 | grep -E "analog diff|adc:|rror" ) 2>&1 | grep -v -E "^$|user|sys"
done
for E in 256 2048 16384; do
  echo "=== timing, $E events"
  ( time TC_EVENTS=4 TC_TIME_EVENTS=$E timeout 300 python scripts/tensor_check.py 2>&1 | grep -E "ms per|rror" ) 2>&1 | grep -v -E "^$|user|sys"
done
