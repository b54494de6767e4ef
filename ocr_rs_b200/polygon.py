"""polygon.rs mirror: clip_polygon / shrink_polygon / expand_polygon (polygon.rs:13-56)."""
import ctypes as C

import numpy as np

from . import _ffi


def expand_polygon(points, factor, ctx=None):
    """points: [(x, y)] -> int32 [m, 2] or None (the reference's Option).  The product form: runs on the device
    (the post-processing's unclip kernel for one polygon)."""
    ctx = ctx if ctx is not None else _ffi.default_context()
    p = np.ascontiguousarray(np.asarray(points, np.int32).reshape(-1, 2))
    cap = 6 * len(p) + 32
    out = np.empty((cap, 2), np.int32)
    n = C.c_int(0)
    _ffi.check(_ffi.lib().ocrb_expand_polygon(ctx.handle, _ffi.ptr(p), len(p), float(factor), _ffi.ptr(out), cap, C.byref(n)))
    return out[: n.value].copy() if n.value > 0 else None


def clip_polygon(points, factor, shrink, return_distance=False):
    """clip_polygon(points, factor, OffsetType::{Shrink, Expand}) (polygon.rs:13-42) on the host — no device needed; the
    same offset / union source the device unclip runs.  -> int32 [m, 2] or None (, signed distance)."""
    p = np.ascontiguousarray(np.asarray(points, np.int32).reshape(-1, 2))
    cap = 6 * len(p) + 32
    out = np.empty((cap, 2), np.int32)
    n, d = C.c_int(0), C.c_double(0.0)
    _ffi.check(_ffi.lib().ocrb_clip_polygon(_ffi.ptr(p), len(p), float(factor), int(bool(shrink)), _ffi.ptr(out), cap, C.byref(n), C.byref(d)))
    res = out[: n.value].copy() if n.value > 0 else None
    return (res, d.value) if return_distance else res


def shrink_polygon(points, factor):
    """shrink_polygon (polygon.rs:44-49): the training-target side of clip_polygon (image_ops.rs:265); host code in the
    reference as well."""
    return clip_polygon(points, factor, True)
