set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 3000 gpurun_out/bench_c4.json; tail -5 gpurun_out/bench_c4.err
python bench.py --steps 5 --warmup 3 --kernel scatter --no-cpu-baseline --no-e2e > gpurun_out/bench_c4_scatter.json 2>&1; tail -c 1500 gpurun_out/bench_c4_scatter.json
python bench.py --steps 5 --warmup 3 --workload dark --no-cpu-baseline > gpurun_out/bench_dark.json 2>&1; tail -c 2500 gpurun_out/bench_dark.json
python bench.py --steps 5 --warmup 3 --workload dark --kernel sweep --no-cpu-baseline --no-e2e > gpurun_out/bench_dark_sweep.json 2>&1; tail -c 1200 gpurun_out/bench_dark_sweep.json
python bench.py --steps 2 --warmup 1 --events-per-step 4096 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c4.csv python bench.py --steps 2 --warmup 1 --events-per-step 4096 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 2 --warmup 1 --events-per-step 4096 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_generate -s 1 -c 1 -o gpurun_out/prof_c4_sweep -f python bench.py --steps 2 --warmup 1 --events-per-step 4096 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
python bench.py --steps 2 --warmup 1 --events-per-step 4096 --workload dark --no-cpu-baseline --no-e2e > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_generate -s 1 -c 1 -o gpurun_out/prof_dark_scatter -f python bench.py --steps 2 --warmup 1 --events-per-step 4096 --workload dark --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_dark.log 2>&1
tail -3 gpurun_out/ncu_full.log gpurun_out/ncu_full_dark.log
ls -la gpurun_out
