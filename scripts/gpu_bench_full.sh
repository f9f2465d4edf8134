# Round-end style measurement: full test suite, smoke, both bench arms, ncu launch list of the same command.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 600 gpurun_out/bench_reference.json
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2800 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --workload dark --no-cpu-baseline > gpurun_out/bench_dark.json 2>&1
python bench.py --no-cpu-baseline > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bench_default.csv python bench.py --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -n 2 gpurun_out/ncu_launch.log | cut -c1-300
