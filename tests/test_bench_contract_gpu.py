"""The bench.py JSON contract (keys the driver reads), on a tiny run."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True,
                         text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_b200_arm_json_line():
    j = _run('--steps', '3', '--warmup', '3', '--events-per-step', '2048', '--no-cpu-baseline')
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'roofline', 'e2e', 'gpu_launches', 'clocks'):
        assert key in j, key
    assert j['unit'] == 'samples/s' and j['higher_is_better'] is True and j['scaling'] == 'weak'
    assert j['n_gpus'] == 1 and j['steps'] == 3 and j['warmup'] == 3 and j['vs_baseline'] is None
    assert 'workload' in j['config'] and 'model' not in j['config']
    r = j['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12
    assert r['achieved'] > 5 and r['traffic'] is not None
    e = j['e2e']
    assert e['value'] > 0 and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 2048 * 1296 * 50 * 2
    assert j['gpu_launches'] == 6 and j['value'] > 1e9            # generation + statistics kernel per step
    assert set(j['config']) == {'workload', 'n_pixels', 'n_bins'}
    assert j['validation']['n_events_reduced'] == 3 * 2048
    assert 300 < j['validation']['mean_adc'] < 500


def test_reference_arm_json_line():
    j = _run('--impl', 'reference', '--steps', '1', '--warmup', '0', '--workload', 'dark')
    assert j['impl'] == 'reference' and j['unit'] == 'samples/s' and j['value'] > 0
    assert j['cpu_baseline']['kind'] == 'port' and j['cpu_baseline']['cores'] >= 1
    assert j['e2e']['value'] == j['value'] and j['e2e']['h2d_bytes_per_step'] == 0
    assert set(j['config']) == {'workload', 'n_pixels', 'n_bins'} and j['config']['n_bins'] == 92     # same keys as the b200 arm
