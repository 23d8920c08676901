"""blu_b200 -- B200-native (sm_100a) implementation of the BLU sparse-LU hot path.

The product is ``libblu_b200.so`` (CUDA kernels + C ABI, ``include/blu_b200.h``); this
package is the thin host-side mirror of the reference crate's ``BLU`` object
(/root/reference/src/blu.rs) over ctypes, plus the synthetic-workload generator.
There is no CPU fallback: without the CUDA library/device every call fails loudly.
"""
from .blu import BLU, BLUBatch, BLUMulti, Status, load_library, library_path  # noqa: F401
from . import gen  # noqa: F401
from .maxvolume import maxvolume  # noqa: F401
