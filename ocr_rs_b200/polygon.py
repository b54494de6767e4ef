"""polygon.rs mirror: expand_polygon (polygon.rs:51-56).  shrink_polygon is training-only
(ground-truth generation) and out of the hot path."""
import ctypes as C

import numpy as np

from . import _ffi


def expand_polygon(points, factor, ctx=None):
    """points: [(x, y)] -> int32 [m, 2] or None (the reference's Option)."""
    ctx = ctx if ctx is not None else _ffi.default_context()
    p = np.ascontiguousarray(np.asarray(points, np.int32).reshape(-1, 2))
    cap = 6 * len(p) + 32
    out = np.empty((cap, 2), np.int32)
    n = C.c_int(0)
    _ffi.check(_ffi.lib().ocrb_expand_polygon(ctx.handle, _ffi.ptr(p), len(p), float(factor), _ffi.ptr(out), cap, C.byref(n)))
    return out[: n.value].copy() if n.value > 0 else None
