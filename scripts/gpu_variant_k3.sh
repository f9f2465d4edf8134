# builds the library with extra nvcc flags ($1) on the GPU box and times K3 alone
cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="$1" > /dev/null 2>&1 && cd ../..
echo "variant: $1"
python scripts/time_stats.py 2>&1 | grep K3
