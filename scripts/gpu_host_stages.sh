# stages of the host pipeline alone and together (diagnosis build); restores the shipped library afterwards
mkdir -p gpurun_out
(cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="-DDCT_HOST_DEBUG" > /dev/null 2>&1)
for S in none copy gen; do DCT_HOST_SKIP=$S python scripts/host_pipeline_stages.py 2>&1 | tail -4; done
(cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make > /dev/null 2>&1)
