# Round-1 measurement set: bench lines (both arms, C4 + dark), ncu launch list of the default bench
# command, one ncu --set full capture of the dominant kernel per workload.  Run under gpurun (1 GPU).
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --workload dark --no-cpu-baseline > gpurun_out/bench_dark.json 2>&1
python bench.py --kernel scatter --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_scatter.json 2>&1
python bench.py --no-cpu-baseline > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bench_default.csv python bench.py --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 2 --warmup 1 --events-per-step 4096 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_generate -s 1 -c 1 -o gpurun_out/prof_c4_sweep -f python bench.py --steps 2 --warmup 1 --events-per-step 4096 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
python bench.py --steps 2 --warmup 1 --events-per-step 4096 --workload dark --no-cpu-baseline --no-e2e > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_generate -s 1 -c 1 -o gpurun_out/prof_dark_scatter -f python bench.py --steps 2 --warmup 1 --events-per-step 4096 --workload dark --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_dark.log 2>&1
python scripts/time_stats.py > gpurun_out/time_stats.txt 2>&1; cat gpurun_out/time_stats.txt
ls gpurun_out
