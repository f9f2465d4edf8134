"""digicamtoy_b200 -- B200-native trace generation for the SST-1M DigiCam toy Monte Carlo.

Drop-in for the ``NTraceGenerator`` path of cta-sst-1m/digicamtoy: same constructor / iterator
API, the per-event Monte Carlo runs in hand-written sm_100a CUDA kernels behind a C ABI
(``include/digicamtoy_b200.h``).  No CPU fallback.
"""
from .tracegenerator import NTraceGenerator  # noqa: F401

__version__ = '0.1.0'
