# ncu --set full of one generation launch: gpu_gen_prof.sh <kernel> <workload> <events> <tag>   (after the same command exited 0 without ncu)
mkdir -p gpurun_out
K=${1:-tensor}; W=${2:-c4}; E=${3:-2048}; T=${4:-$K}
python scripts/run_gen_once.py $K $W $E > gpurun_out/plain_gen_$T.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_generate -s 1 -c 1 -o gpurun_out/prof_gen_$T -f python scripts/run_gen_once.py $K $W $E > gpurun_out/ncu_gen_$T.log 2>&1
tail -n 2 gpurun_out/plain_gen_$T.log gpurun_out/ncu_gen_$T.log
