"""Two K1c launches on the C4 geometry with sub_binning=1 (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from digicamtoy_b200 import NTraceGenerator, workloads
g = NTraceGenerator(seed=1, sub_binning=1, **workloads.c4_full_camera())
assert g.kernel == 'grid'
g.generate(4096); g.generate(4096, first_event=4096)
torch.cuda.synchronize()
print('ok')
