#!/usr/bin/env python
"""Builds profiles/<round>/SUMMARY.md from the bench JSON lines a gpurun call left in gpurun_out/
(copies the JSON lines next to it so the numbers are traceable).  Usage: make_profile_summary.py r01"""
import glob, json, os, shutil, sys
rnd = sys.argv[1] if len(sys.argv) > 1 else 'r01'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, 'profiles', rnd)
os.makedirs(out_dir, exist_ok=True)
rows = []
# bench lines already curated into profiles/<round>/ win; anything else comes from gpurun_out/
paths = {os.path.basename(p_): p_ for p_ in glob.glob(os.path.join(root, 'gpurun_out', 'bench_*.json'))}
paths.update({os.path.basename(p_): p_ for p_ in glob.glob(os.path.join(out_dir, 'bench_*.json'))})
for name_, path in sorted(paths.items()):
    try:
        line = [l for l in open(path).read().strip().splitlines() if l.startswith('{')][-1]
        j = json.loads(line)
    except Exception:
        continue
    if os.path.dirname(path) != out_dir:
        shutil.copy(path, os.path.join(out_dir, name_))
    rows.append((name_, j))
peaks = json.load(open(os.path.join(root, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(root, 'MEASURED_PEAKS.json')) else {}
with open(os.path.join(out_dir, 'SUMMARY.md'), 'w') as f:
    f.write('# Measured on B200 (%s)\n\n' % rnd)
    f.write('HBM roofline denominator: %s GB/s (MEASURED_PEAKS.json, "of measured").\n\n' % peaks.get('hbm_gbs', '6650 (fallback)'))
    f.write('| file | impl | workload | kernel | GPUs | samples/s | events/s | ms/step | K1 ms | K1 GB/s | frac of HBM | e2e samples/s | CPU oracle samples/s (cores) | SM clk MHz | reasons |\n')
    f.write('|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n')
    for name, j in rows:
        r = j.get('roofline') or {}
        c = j.get('cpu_baseline') or {}
        e = j.get('e2e') or {}
        k = j.get('clocks') or {}
        cfg = j.get('config', {})
        f.write('| %s | %s | %s | %s | %s | %.3e | %s | %.2f | %s | %s | %s | %s | %s | %s | %s |\n' % (
            name, j.get('impl', 'b200'), cfg.get('workload', '')[:40], (j.get('run_details') or cfg).get('kernel', '-'), j.get('n_gpus'),
            j['value'], ('%.3e' % j['events_per_s']) if 'events_per_s' in j else '-', j['ms_per_step'],
            ('%.2f' % r['kernel_ms']) if r else '-', ('%.1f' % r['achieved']) if r else '-',
            ('%.4f' % r['frac']) if r else '-', ('%.3e' % e['value']) if e else '-',
            ('%.3e (%d)' % (c['value'], c['cores'])) if c else '-', k.get('sm_mhz', '-'), ','.join(k.get('reasons', [])) or 'none'))
    f.write('\nRows with K1 = 38.2-38.3 ms (bench_c4_a, bench_c5_full_1e7_n8) were taken with K1t as a CTA kernel, '
            'the others with the persistent kernel that replaced it later in the round (36.8-37.1 ms).\n')
    f.write('\nncu summaries in this directory: ' + ', '.join(sorted(os.path.basename(p) for p in glob.glob(os.path.join(out_dir, 'ncu_*.txt')))) + '\n')
    f.write('\nLaunch lists: ' + ', '.join(sorted(os.path.basename(p) for p in glob.glob(os.path.join(out_dir, 'launches_*.csv')))) + '\n')
import csv, collections
lp = os.path.join(out_dir, 'launches_bench_default.csv')
if os.path.exists(lp):
    rows_ = list(csv.reader(open(lp)))
    hi = next(i for i, r in enumerate(rows_) if r and r[0] == 'ID')
    hdr = rows_[hi]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows_[hi + 1:]:
        if len(r) < len(hdr):
            continue
        scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ix['Metric Unit']], 1e-6)
        a_ = agg[r[ix['Kernel Name']][:70]]
        a_[0] += 1; a_[1] += float(r[ix['Metric Value']].replace(',', '')) * scale
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, 'SUMMARY.md'), 'a') as f:
        f.write('\n## ncu launch list of `python bench.py --no-cpu-baseline` (first 200 launches, `launches_bench_default.csv`)\n\n')
        ev = next((j for n_, j in rows if n_ == 'bench_n1.json'), None)
        f.write('Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n')
        if ev:
            f.write('In the CUDA-event timing of the same command K1 is %.1f of %.1f ms per step (%.0f %%).\n' % (
                ev['roofline']['kernel_ms'], ev['ms_per_step'], 100 * ev['roofline']['kernel_ms'] / ev['ms_per_step']))
        f.write('(Launches shorter than a full step are the chunks of the end-to-end leg and the iterator batches.)\n\n')
        f.write('| kernel | launches | total ms | share |\n|---|---|---|---|\n')
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.2f | %.1f %% |\n' % (k, v[0], v[1], 100 * v[1] / tot))
print(open(os.path.join(out_dir, 'SUMMARY.md')).read())
