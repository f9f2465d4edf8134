"""One K3 launch on a dark-run cube (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads
from digicamtoy_b200.stats import CubeStats
which = sys.argv[1] if len(sys.argv) > 1 else 'dark'
kw = workloads.c1_dark(0.) if which == 'dark' else workloads.c4_full_camera()
g = NTraceGenerator(seed=1, **kw)
cube = g.generate(4096)
st = CubeStats(g)
st.accumulate(cube); st.accumulate(cube)
torch.cuda.synchronize()
print('ok')
