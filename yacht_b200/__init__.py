"""yacht_b200 -- B200-native (sm_100a) implementation of YACHT's data-parallel hot path.

Layout: ``csrc/`` holds the CUDA kernels and the C ABI (``include/yacht_gpu.h``) plus the host C++
drop-in for the reference's ``run_yacht_train_core`` executable; the Python modules mirror the
reference's own operator interface for the path (``utils.run_yacht_train_core``,
``hypothesis_recovery_src.*``) on top of that ABI.
"""
__version__ = "0.1.0"
