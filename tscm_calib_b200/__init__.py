"""B200-native TSCM-Calib calibration solve (TS projection/Jacobian + LM/Schur).

The product is `libtscm_b200.so` (hand-written sm_100a FP64 CUDA behind the C-ABI
of include/tscm.h).  This package is the Python-side plumbing used by the tests
and bench.py: ctypes bindings (`capi`) and the synthetic-problem generator
(`synth`).  Importing the package does not load CUDA; `capi.load_library()` does
and raises if the library is missing — there is no CPU fallback.
"""
from . import capi, synth  # noqa: F401

__all__ = ["capi", "synth"]
__version__ = "0.1.0"
