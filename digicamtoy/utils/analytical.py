"""Import-path shim for ``digicamtoy.utils.analytical`` (reference digicamtoy/utils/analytical.py)."""
from digicamtoy_b200.params import gain_drop, true_gain_drop, true_pde_drop, true_xt_drop  # noqa: F401
