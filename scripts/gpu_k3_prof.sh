# ncu --set full of one K3 launch per workload (after the same command exited 0 without ncu)
mkdir -p gpurun_out
for w in c4 dark; do
  python scripts/run_k3_once.py $w > gpurun_out/plain_k3_$w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_stats -s 1 -c 1 -o gpurun_out/prof_k3_$w -f python scripts/run_k3_once.py $w > gpurun_out/ncu_k3_$w.log 2>&1
  tail -n 2 gpurun_out/ncu_k3_$w.log
done
