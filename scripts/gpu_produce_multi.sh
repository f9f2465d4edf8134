# f-4: the (nsb, n_photon) levels of a YAML dealt round-robin to torchrun ranks
rm -rf /tmp/pd_out && mkdir -p /tmp/pd_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 produce_data.py -y config_files/ac_dc.yml --output_directory /tmp/pd_out --events_per_level 10 --seed 1 2>&1 | grep -c "File saved"
ls /tmp/pd_out | wc -l
python - <<'PY'
import numpy as np, glob
files = sorted(glob.glob('/tmp/pd_out/*'), key=lambda p: int(p.split('_')[-1].split('.')[0]))
for i, f in enumerate(files):
    z = np.load(f)
    print(i, f.split('/')[-1], float(z['config/nsb_rate']), float(z['config/n_photon']), z['data/adc_count'].shape, round(float(z['data/adc_count'].mean()), 2), round(float(z['data/true_baseline'].mean()), 2))
PY
