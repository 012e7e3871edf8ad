"""wfcrl_b200: B200-native batched Floris backend + env shells for ifpen/wfcrl-env's ``*_Floris`` environments.

Compute lives in ``libwfcrl_b200.so`` (hand-written sm_100a CUDA kernels behind the C-ABI of
``include/wfcrl_b200.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from .layouts import get_layout, layout_xy, named_layouts  # noqa: F401
