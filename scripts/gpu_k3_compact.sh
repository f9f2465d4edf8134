# K3: compact tile layout (one tile buffer, windowed CTA histogram) against the other layouts.
# $1 = "long" (default): long traces; "short": would the 50-bin cube gain from it too?
set -u
mkdir -p gpurun_out
python -m pytest tests/test_generate_gpu.py -q -x -k "stats" 2>&1 | tail -5
if [ "${1:-long}" = "long" ]; then   # (the TILE_MAX_BINS switch and the warp kernel it selected are gone)
  set -- "" "-DDCT_STATS_COMPACT_SINGLE=0" "-DDCT_STATS_COMPACT_HIST=4095" "-DDCT_STATS_COMPACT_HIST=512" "-DDCT_STATS_COMPACT_HIST=2048"
else
  set -- "-DDCT_STATS_COMPACT_BINS=64" "-DDCT_STATS_COMPACT_BINS=64 -DDCT_STATS_COMPACT_HIST=2048" "-DDCT_STATS_COMPACT_BINS=64 -DDCT_STATS_COMPACT_SINGLE=0" "-DDCT_STATS_COMPACT_BINS=64 -DDCT_STATS_COMPACT_HIST=4095"
fi
for v in "$@"; do
  bash scripts/gpu_variant_k3.sh "$v"
done 2>&1 | tee -a gpurun_out/k3_compact.txt
