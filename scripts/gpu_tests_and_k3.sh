mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/time_stats.py
