mkdir -p gpurun_out
python scripts/run_k3_once.py c4 > gpurun_out/plain_k3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_stats -s 1 -c 1 -o gpurun_out/prof_k3_c4 -f python scripts/run_k3_once.py c4 > gpurun_out/ncu_k3.log 2>&1
tail -n 2 gpurun_out/ncu_k3.log
