#!/usr/bin/env python
"""Builds profiles/<round>/SUMMARY.md from the bench JSON lines a gpurun call left in gpurun_out/
(copies the JSON lines next to it so the numbers are traceable).  Usage: make_profile_summary.py r01"""
import glob, json, os, shutil, sys
rnd = sys.argv[1] if len(sys.argv) > 1 else 'r01'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, 'profiles', rnd)
os.makedirs(out_dir, exist_ok=True)
rows = []
for path in sorted(glob.glob(os.path.join(root, 'gpurun_out', 'bench_*.json'))):
    try:
        line = [l for l in open(path).read().strip().splitlines() if l.startswith('{')][-1]
        j = json.loads(line)
    except Exception:
        continue
    shutil.copy(path, os.path.join(out_dir, os.path.basename(path)))
    rows.append((os.path.basename(path), j))
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
            name, j.get('impl', 'b200'), cfg.get('workload', '')[:40], cfg.get('kernel', '-'), j.get('n_gpus'),
            j['value'], ('%.3e' % j['events_per_s']) if 'events_per_s' in j else '-', j['ms_per_step'],
            ('%.2f' % r['kernel_ms']) if r else '-', ('%.1f' % r['achieved']) if r else '-',
            ('%.4f' % r['frac']) if r else '-', ('%.3e' % e['value']) if e else '-',
            ('%.3e (%d)' % (c['value'], c['cores'])) if c else '-', k.get('sm_mhz', '-'), ','.join(k.get('reasons', [])) or 'none'))
    f.write('\nncu summaries in this directory: ' + ', '.join(sorted(os.path.basename(p) for p in glob.glob(os.path.join(out_dir, 'ncu_*.txt')))) + '\n')
    f.write('\nLaunch lists: ' + ', '.join(sorted(os.path.basename(p) for p in glob.glob(os.path.join(out_dir, 'launches_*.csv')))) + '\n')
print(open(os.path.join(out_dir, 'SUMMARY.md')).read())
