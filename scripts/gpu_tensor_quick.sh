# tensor-core kernel: parity against the sweep kernel (64 events) + timing (16384 events), bounded
mkdir -p gpurun_out
TC_EVENTS=64 TC_TIME_EVENTS=16384 timeout 300 python scripts/tensor_check.py 2>&1 | grep -E "analog diff|adc:|ms per|rror"
