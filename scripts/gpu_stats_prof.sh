# ncu --set full of one K3 (statistics) launch of the bench: gpu_stats_prof.sh <workload> <tag>
mkdir -p gpurun_out
W=${1:-c4}; T=${2:-stats_$W}
ncu --set full --clock-control none --import-source on -k regex:k_stats -s 2 -c 1 -o gpurun_out/prof_$T -f python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$T.log 2>&1
tail -n 2 gpurun_out/ncu_$T.log
