"""peba1_b200 — B200-native TFHE gate-bootstrapping engine for the PEBA1 face-matching circuits.

The product is native code: ``csrc/libpeba1_cuda.so`` (hand-written sm_100a kernels behind the C ABI of
``include/peba1_cuda.h``) and ``host/libpeba1_tfhe.so`` (the TFHE ``boots*`` gate API and PEBA1's ``Math.h``
routines as a recording, level-batching scheduler).  This Python package is only a thin ``ctypes`` view used by
the tests and the benchmark; it never computes anything itself and never imports ``oracle/``.
"""
from .cuda import (Engine, Params, Gate, OP, PEBA1Error, lib_path, load_library, default_params,  # noqa: F401
                   DECLARED_SYMBOLS)
