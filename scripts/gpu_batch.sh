# f-4: batch launcher test + a timed slice of the commissioning grid (40 files x 1000 events, 1296 px)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_produce_and_shards_gpu.py -q -x 2>&1 | tail -4
rm -rf /dev/shm/mc && python batch_produce.py --ac 100:103 --file_id 0:9 --output_directory /dev/shm/mc --seed 1 --writers ${WRITERS:-16} --max_in_flight ${INFLIGHT:-20} 2>&1 | tail -3 | tee gpurun_out/batch_produce.txt
du -sh /dev/shm/mc | tee -a gpurun_out/batch_produce.txt; ls /dev/shm/mc | head -3; nproc; free -g | head -2
rm -rf /dev/shm/mc
