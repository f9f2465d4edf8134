# ncu --set full of one K3 launch (C4 cube) for a library built with extra flags ($1)
cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="$1" > /dev/null 2>&1 && cd ../..
mkdir -p gpurun_out
python scripts/run_k3_once.py c4 > gpurun_out/plain_k3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_stats -s 1 -c 1 -o gpurun_out/prof_k3_variant -f python scripts/run_k3_once.py c4 > gpurun_out/ncu_k3.log 2>&1
tail -n 1 gpurun_out/ncu_k3.log
