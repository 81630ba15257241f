"""dbsp_b200 — B200-native Z-set delta hot path of DBSP.

Host-side mirror of the reference's Stream/operator API over the C ABI in
include/dbsp_b200.h (CUDA library `libdbsp_b200.so`, sm_100a).  Importable as
`dbsp_b200` through the shim at the repository root.
"""
from . import _capi as capi
from .circuit import FoldCount, FoldSum, Max, Min, RootCircuit, Stream
from .zset import Backend, Batch, Batcher, Merger, Proj, Schema, Spine, col, const, key, lval, rval, val

__all__ = [
    "capi", "RootCircuit", "Stream", "Max", "Min", "FoldCount", "FoldSum", "Backend", "Batch", "Proj",
    "Schema", "Spine", "Batcher", "Merger", "key", "lval", "rval", "val", "col", "const",
]
