"""fries_b200 -- B200-native FRI hot path.  Python here is only the thin host mirror used by
the tests and bench.py; the product is libfries_b200.so (C-ABI in include/fries_b200.h) and the C++ host
layer in fries_b200/host/."""
from . import _capi  # noqa: F401  (raises ImportError when the CUDA library has not been built)
from .api import Context, Mol, Vec, find_preserve, sys_comp, comp_sub, hash_owner, bit_op, hh_batch  # noqa: F401
from .api import piv_samp_serial, adjust_probs, piv_budget, piv_comp  # noqa: F401
