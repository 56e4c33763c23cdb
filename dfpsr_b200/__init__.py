"""dfpsr_b200 — B200-native (sm_100a) rendering hot path behind DFPSR's API.

The product is the C-ABI shared library built from dfpsr_b200/csrc (include/dfpsr_b200.h) plus the C++14
`dsr::` shim in dfpsr_b200/host. This Python package is the thin host-side harness used by tests and
bench.py: ctypes bindings (lib.py), POD mirrors (abi.py) and synthetic scene generators (scenes.py).
There is no CPU fallback: importing dfpsr_b200.lib and calling a compute entry point without the CUDA
library or without a GPU raises.
"""
