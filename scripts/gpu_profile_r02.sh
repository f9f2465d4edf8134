# Round-2 evidence set (1 GPU): GPU tests, smoke, both bench arms, launch list of the default bench
# command (ncu, serialised and cold-cache: shares, not absolutes), one ncu --set full capture of K1t.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --no-cpu-baseline > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bench_default.csv python bench.py --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
bash scripts/gpu_gen_prof.sh tensor c4 2048 tensor_final
bash scripts/gpu_full_workloads.sh
python scripts/sass_evidence.py gpurun_out/sass_evidence.txt
cut -c1-600 gpurun_out/bench_n1.json; cut -c1-400 gpurun_out/bench_reference.json
