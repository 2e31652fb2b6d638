"""fmftsaxs-b200: B200-native FFT-SAXS dimer scoring (the `correlate` path of libfmftsaxs).

The product is the C shared library ``libfmftsaxs.so`` (C11 host code + sm_100a CUDA kernels) built
in-tree by ``libfmftsaxs_b200.build``; this package is only its ctypes binding.  There is no CPU
implementation: loading fails loudly when the library has not been built, and every compute entry
point fails when no CUDA device is visible.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
