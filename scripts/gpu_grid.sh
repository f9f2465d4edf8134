# K1c (on-grid NSB kernel): parity tests, guards, timing against the other modes; variants if $1 = variants
mkdir -p gpurun_out
python -m pytest tests/test_generate_gpu.py tests/test_errors_gpu.py tests/test_guards_gpu.py tests/test_replay_parity_gpu.py -q -x 2>&1 | tail -3
python scripts/time_modes.py 2>&1 | tee gpurun_out/modes_timing.txt
if [ "${1:-}" = "variants" ]; then
  for v in "-DDCT_GRID_MIN_BLOCKS=4" "-DDCT_GRID_MIN_BLOCKS=6" "-DDCT_GRID_MIN_BLOCKS=8"; do
    (cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="$v" > /dev/null 2>&1)
    echo "variant: $v"; python scripts/time_modes.py 2>&1 | grep sub_binning
  done | tee gpurun_out/grid_variants.txt
fi
