# the tail-stop of the inversion samplers: test with the shipped build, then show the failure without it
mkdir -p gpurun_out
python -m pytest tests/test_generate_gpu.py -q -x -k "largest_uniforms or sampler or branching or identical" 2>&1 | tail -3
(cd digicamtoy_b200/csrc && rm -f ../libdigicamtoy_sm100.so && make EXTRA="-DDCT_EXP_NO_TAIL_STOP=1" > /dev/null 2>&1)
python - <<'PY' 2>&1 | tee gpurun_out/tail_stop.txt
import numpy as np, torch
from digicamtoy_b200 import native
def sample(kind, a, n=64):
    out = torch.empty(n, dtype=torch.float32, device='cuda')
    native.test_sample(kind, a, 0, 11, n, out.data_ptr()); torch.cuda.synchronize()
    return out.cpu().numpy()
print('WITHOUT the tail stop (-DDCT_EXP_NO_TAIL_STOP=1): result for the largest uniforms 1-2^-23, 1-2*2^-23, ...')
bad = 0
lams = np.linspace(0.004, 11.99, 300)
for lam in lams:
    v = sample('poisson_inv_top', lam)
    if v.max() >= 200: bad += 1
print('poisson_inv: %d of %d means return the cap (200) for at least one of the 64 largest uniforms' % (bad, len(lams)))
for lam in (0.08, 0.32, 4.0):
    print('  lambda %.2f ->' % lam, sample('poisson_inv_top', lam)[:6])
bad = 0
xts = np.linspace(0.005, 0.5, 100)
for xt in xts:
    if sample('borel_top', xt).max() >= 256: bad += 1
print('borel: %d of %d crosstalk values return the cap (256)' % (bad, len(xts)))
print('  xt 0.08 ->', sample('borel_top', 0.08)[:6])
PY
