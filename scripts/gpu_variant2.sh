# kernel time of named bench workloads under build variants: gpu_variant2.sh "<workloads>" "<EXTRA 1>" "<EXTRA 2>" ...
mkdir -p gpurun_out
W="$1"; shift
for V in "$@"; do
  (cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="$V" > /dev/null 2>&1)
  echo "=== variant: $V"
  for w in $W; do
    k=auto; wl=$w; case $w in c4sweep) wl=c4; k=sweep;; c4tensor) wl=c4; k=tensor;; esac
    python bench.py --workload $wl --kernel $k --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -o '"kernel_ms": [0-9.]*' | sed "s/^/$w /"
  done
done
