"""arcs_b200 -- B200-native (sm_100a) ARKS hot path of bcgsc/arcs behind a C ABI.

The product is `lib/libarks_b200.so` (hand-written CUDA, see csrc/) and the `arcs-b200`
C++ command line (host/).  This package is the thin Python (ctypes) mirror of
include/arks_b200.h used by tests/ and bench.py.  There is no CPU fallback: if the
library is missing, or there is no CUDA device, calls raise.
"""
from .api import (  # noqa: F401
    ArksError,
    ArksIndex,
    IndexStats,
    MapStats,
    comm_init_local,
    comm_unique_id,
    head_tail_table,
    lib_path,
    load_library,
    merge_pmap_local,
)

__all__ = ["ArksError", "ArksIndex", "IndexStats", "MapStats", "head_tail_table", "lib_path", "load_library",
           "comm_unique_id", "comm_init_local", "merge_pmap_local"]
