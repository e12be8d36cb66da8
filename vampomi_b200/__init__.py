"""vampomi_b200 — B200 (sm_100a) implementation of gVAMPomi's VAMP inference hot path.

The product is the shared library built from ``csrc/`` (CUDA kernels + C ABI + C++ host driver, headers in
``include/``) and the ``main_meth`` executable; this package only binds it for tests and benchmarks.
"""
from .capi import (Shard, Solver, VampomiError, comm_unique_id, device_count, divide_work, exported_symbols,  # noqa: F401
                   load_library, main)
