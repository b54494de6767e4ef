"""ctypes binding of include/ocrb.h (libocrb.so, built in-tree by ocr_rs_b200/csrc/Makefile).

This is the same C ABI a Rust `ocrb-sys` crate would bind.  There is no
CPU fallback: if the shared library is missing, importing a compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OCRB_LIB_PATH") or os.path.join(_HERE, "libocrb.so")  # (override: A/B runs of two builds)

OK = 0
MODE_FP32, MODE_BF16 = 0, 1
U8, F32 = 0, 1
PIXELS_RGBA, PIXELS_LUMA = 0, 1

c_p = C.c_void_p
i64 = C.c_int64


class MetricsItem(C.Structure):
    _fields_ = [("precision", C.c_double), ("recall", C.c_double), ("hmean", C.c_double),
                ("gt_care", C.c_int64), ("det_care", C.c_int64), ("det_matched", C.c_int64)]


class PostprocParams(C.Structure):
    _fields_ = [("thresh", C.c_double), ("box_thresh", C.c_double), ("min_size", C.c_double),
                ("unclip_factor", C.c_double)]


# name -> (restype, argtypes); every symbol include/ocrb.h declares
SIGNATURES = {
    "ocrb_version": (C.c_int, []),
    "ocrb_last_error": (C.c_char_p, []),
    "ocrb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ocrb_ctx_create": (C.c_int, [C.c_int, C.POINTER(c_p)]),
    "ocrb_ctx_destroy": (C.c_int, [c_p]),
    "ocrb_ctx_synchronize": (C.c_int, [c_p]),
    "ocrb_ctx_stream": (c_p, [c_p]),
    "ocrb_ctx_wait_stream": (C.c_int, [c_p, c_p]),
    "ocrb_ctx_device": (C.c_int, [c_p]),
    "ocrb_ctx_launch_count": (i64, [c_p]),
    "ocrb_ctx_profile_begin": (C.c_int, [c_p]),
    "ocrb_ctx_profile_end": (C.c_int, [c_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ocrb_resize_dims": (C.c_int, [C.c_int] * 4 + [C.POINTER(C.c_int)] * 2),
    "ocrb_preprocess_rgba": (C.c_int, [c_p, c_p, C.c_int, C.c_int, C.c_int, C.c_int, c_p,
                                       C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ocrb_preprocess_rgba_batch": (C.c_int, [c_p, c_p, c_p, c_p, c_p, C.c_int, C.c_int, C.c_int, c_p, c_p]),
    "ocrb_image_info": (C.c_int, [c_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ocrb_debug_decode_host": (C.c_int, [c_p, C.c_size_t, c_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ocrb_decode_images": (C.c_int, [c_p, C.POINTER(c_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, c_p, c_p]),
    "ocrb_preprocess_files": (C.c_int, [c_p, C.POINTER(c_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, c_p, c_p]),
    "ocrb_convert_image_to_tensor": (C.c_int, [c_p, c_p, i64, c_p]),
    "ocrb_convert_tensor_to_image": (C.c_int, [c_p, c_p, i64, C.c_float, c_p]),
    "ocrb_load_image_as_tensor": (C.c_int, [c_p, c_p, i64, c_p]),
    "ocrb_det_create": (C.c_int, [c_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(c_p), C.POINTER(i64), C.c_int,
                                  C.POINTER(c_p)]),
    "ocrb_det_destroy": (C.c_int, [c_p]),
    "ocrb_det_forward": (C.c_int, [c_p, c_p, C.c_int, C.c_int, C.c_int, C.c_int, c_p]),
    "ocrb_det_tap": (C.c_int, [c_p, C.c_char_p, c_p, i64]),
    "ocrb_binarize": (C.c_int, [c_p, c_p, i64, C.c_double, c_p]),
    "ocrb_box_score_fast": (C.c_int, [c_p, c_p, C.c_int, C.c_int, c_p, C.c_int, C.POINTER(C.c_double)]),
    "ocrb_min_area_bounding_box": (C.c_int, [c_p, c_p, C.c_int, c_p, C.POINTER(C.c_double)]),
    "ocrb_expand_polygon": (C.c_int, [c_p, c_p, C.c_int, C.c_double, c_p, C.c_int, C.POINTER(C.c_int)]),
    "ocrb_clip_polygon": (C.c_int, [c_p, C.c_int, C.c_double, C.c_int, c_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "ocrb_postproc_default_params": (None, [C.POINTER(PostprocParams)]),
    "ocrb_get_boxes_and_box_scores": (C.c_int, [c_p, c_p, c_p, C.c_int, C.c_int, C.c_int, C.POINTER(PostprocParams),
                                                C.POINTER(c_p)]),
    "ocrb_get_polygons_from_bitmap": (C.c_int, [c_p, c_p, c_p, c_p, C.c_int, C.c_int, C.POINTER(PostprocParams),
                                                C.POINTER(c_p)]),
    "ocrb_polygons_num_images": (C.c_int, [c_p]),
    "ocrb_polygons_image_offsets": (C.POINTER(i64), [c_p]),
    "ocrb_polygons_point_offsets": (C.POINTER(i64), [c_p]),
    "ocrb_polygons_xy": (C.POINTER(C.c_uint32), [c_p]),
    "ocrb_polygons_scores": (C.POINTER(C.c_double), [c_p]),
    "ocrb_polygons_stats": (C.POINTER(i64), [c_p]),
    "ocrb_polygons_free": (None, [c_p]),
    "ocrb_ccl_labels": (C.c_int, [c_p, c_p, C.c_int, C.c_int, C.c_int, c_p, c_p]),
    "ocrb_debug_conv_geometry": (C.c_int, [C.c_int, C.c_int, C.c_int, c_p]),
    "ocrb_debug_approx_polygon_host": (C.c_int, [c_p, C.c_int64, c_p, C.c_int64, C.POINTER(C.c_int64)]),
    "ocrb_debug_min_area_bounding_box_host": (C.c_int, [c_p, C.c_int, c_p, C.POINTER(C.c_double)]),
    "ocrb_debug_pipeline_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_p, c_p, C.c_int, c_p]),
    "ocrb_find_contours": (C.c_int, [c_p, c_p, C.c_int, C.c_int, c_p, c_p, i64, c_p, i64, C.POINTER(i64),
                                     C.POINTER(i64)]),
    "ocrb_approx_polygon": (C.c_int, [c_p, c_p, i64, c_p, i64, C.POINTER(i64)]),
    "ocrb_rec_create": (C.c_int, [c_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(c_p), C.POINTER(i64), C.POINTER(c_p)]),
    "ocrb_rec_destroy": (C.c_int, [c_p]),
    "ocrb_rec_forward": (C.c_int, [c_p, c_p, C.c_int, c_p, c_p, c_p]),
    "ocrb_rec_forward_u8": (C.c_int, [c_p, c_p, C.c_int, c_p, c_p, c_p]),
    "ocrb_class_to_char": (C.c_char, [C.c_int]),
    "ocrb_varstore_open": (C.c_int, [C.c_char_p, C.POINTER(c_p)]),
    "ocrb_varstore_count": (C.c_int, [c_p]),
    "ocrb_varstore_name": (C.c_char_p, [c_p, C.c_int]),
    "ocrb_varstore_tensor": (C.c_int, [c_p, C.c_int, C.POINTER(C.POINTER(C.c_float)), C.POINTER(i64), C.POINTER(C.POINTER(i64)), C.POINTER(C.c_int)]),
    "ocrb_varstore_close": (None, [c_p]),
    "ocrb_det_create_from_file": (C.c_int, [c_p, C.c_char_p, C.c_int, C.POINTER(c_p)]),
    "ocrb_rec_create_from_file": (C.c_int, [c_p, C.c_char_p, C.POINTER(c_p)]),
    "ocrb_detect_and_recognize": (C.c_int, [c_p, c_p, c_p, c_p, C.c_int, C.c_int, C.c_int, C.POINTER(PostprocParams),
                                            c_p, C.c_int, c_p, C.POINTER(c_p)]),
    "ocrb_detect_and_read": (C.c_int, [c_p, c_p, c_p, c_p, C.c_int, C.c_int, C.c_int, C.POINTER(PostprocParams), C.c_int, C.POINTER(c_p)]),
    "ocrb_polygons_glyphs_per_polygon": (C.c_int, [c_p]),
    "ocrb_polygons_glyph_classes": (C.POINTER(C.c_int32), [c_p]),
    "ocrb_crop_glyphs": (C.c_int, [c_p, c_p, C.c_int, C.c_int, c_p, C.c_int, C.c_int, c_p]),
    "ocrb_detect_and_read_sharded": (C.c_int, [c_p, c_p, c_p, C.c_int, C.c_int, C.c_int, C.POINTER(PostprocParams), C.c_int, C.POINTER(c_p)]),
    "ocrb_polygon_iou": (C.c_int, [c_p, C.c_int, c_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ocrb_evaluate_image": (C.c_int, [c_p, c_p, C.c_int, c_p, c_p, c_p, C.c_int, C.POINTER(MetricsItem)]),
    "ocrb_combine_results": (C.c_int, [C.POINTER(MetricsItem), C.c_int] + [C.POINTER(C.c_double)] * 3),
    "ocrb_shards_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(c_p), C.POINTER(i64), C.c_int,
                                     C.c_int, C.POINTER(C.c_char_p), C.POINTER(c_p), C.POINTER(i64), C.POINTER(c_p)]),
    "ocrb_shards_create_from_files": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(c_p)]),
    "ocrb_shards_destroy": (C.c_int, [c_p]),
    "ocrb_shards_count": (C.c_int, [c_p]),
    "ocrb_shards_device": (C.c_int, [c_p, C.c_int]),
    "ocrb_shards_launch_count": (i64, [c_p]),
    "ocrb_shard_range": (C.c_int, [i64, C.c_int, C.c_int, C.POINTER(i64), C.POINTER(i64)]),
    "ocrb_detect_and_recognize_sharded": (C.c_int, [c_p, c_p, c_p, C.c_int, C.c_int, C.c_int, C.POINTER(PostprocParams),
                                                    c_p, C.c_int, c_p, C.POINTER(c_p)]),
    "ocrb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(c_p)]),
    "ocrb_host_free": (C.c_int, [c_p]),
}

_lib = None


class OcrbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libocrb error {code}: {msg}")
        self.code = code


def lib():
    """Loads libocrb.so; raises (loudly) if it has not been built — no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(libocrb has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise OcrbError(rc, lib().ocrb_last_error().decode("utf-8", "replace"))


def ptr(x):
    """void* of a numpy array, a torch tensor (host or cuda) or a raw int address."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"], "array must be C-contiguous"
        return x.ctypes.data
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous(), "tensor must be contiguous"
        if x.is_cuda:
            # ocrb.h stream-ordering contract: a device input must be complete before the call.  The tensor may
            # still be in flight on torch's current stream (and libocrb launches on its own stream), so its
            # producer is synchronised here; outputs are complete when the call returns.
            import torch
            torch.cuda.current_stream(x.device).synchronize()
        return x.data_ptr()
    raise TypeError(type(x))


class Context:
    """ocrb_ctx: one per (device, stream) and host thread (replaces main.rs:26-28 DEVICE)."""

    def __init__(self, device=0):
        self._h = c_p()
        check(lib().ocrb_ctx_create(int(device), C.byref(self._h)))
        self.device = device

    @property
    def handle(self):
        return self._h

    def synchronize(self):
        check(lib().ocrb_ctx_synchronize(self._h))

    @property
    def stream(self):
        return lib().ocrb_ctx_stream(self._h)

    @property
    def launch_count(self):
        return int(lib().ocrb_ctx_launch_count(self._h))

    def profile_begin(self):
        check(lib().ocrb_ctx_profile_begin(self._h))

    def profile_end(self):
        """-> {kernel name: (launch count, total ms)} since profile_begin()."""
        need = C.c_size_t(0)
        check(lib().ocrb_ctx_profile_end(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib().ocrb_ctx_profile_end(self._h, buf, need.value, C.byref(need)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(" ", 2)
            out[name] = (int(cnt), float(ms))
        return out

    def close(self):
        if self._h:
            lib().ocrb_ctx_destroy(self._h)
            self._h = c_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def weights_to_c(weights):
    """dict name -> float32 array  =>  (n, names[], data[], numel[], keepalive)."""
    names = list(weights.keys())
    arrs = [np.ascontiguousarray(np.asarray(weights[k], dtype=np.float32)) for k in names]
    n = len(names)
    c_names = (C.c_char_p * n)(*[k.encode() for k in names])
    c_data = (c_p * n)(*[a.ctypes.data for a in arrs])
    c_numel = (i64 * n)(*[a.size for a in arrs])
    return n, c_names, c_data, c_numel, arrs


class Polygons:
    """PolygonScores (metrics.rs:32-35) backed by an ocrb_polygons handle.  The flat arrays
    (image_offsets, point_offsets, xy, all_scores, stats) are copied out eagerly; the
    per-image Python lists `polygons` / `scores` are built on first use."""

    def __init__(self, handle=None, arrays=None):
        if arrays is not None:
            self.image_offsets, self.point_offsets, self.xy, self.all_scores, self.stats = arrays
        else:
            L = lib()
            nb = L.ocrb_polygons_num_images(handle)
            io = np.ctypeslib.as_array(L.ocrb_polygons_image_offsets(handle), shape=(nb + 1,)).copy()
            npoly = int(io[-1])
            po = np.ctypeslib.as_array(L.ocrb_polygons_point_offsets(handle), shape=(npoly + 1,)).copy()
            npts = int(po[-1])
            # an empty result has no backing storage (NULL data pointers)
            xy = (np.ctypeslib.as_array(L.ocrb_polygons_xy(handle), shape=(npts * 2,)).copy().reshape(-1, 2)
                  if npts > 0 else np.zeros((0, 2), np.uint32))
            sc = np.ctypeslib.as_array(L.ocrb_polygons_scores(handle), shape=(npoly,)).copy() if npoly > 0 else np.zeros(0, np.float64)
            self.stats = np.ctypeslib.as_array(L.ocrb_polygons_stats(handle), shape=(nb * 5,)).copy().reshape(nb, 5)
            k = L.ocrb_polygons_glyphs_per_polygon(handle)
            if k > 0:  # ocrb_detect_and_read: classes of the glyph tiles cut from every polygon
                self.glyph_classes = (np.ctypeslib.as_array(L.ocrb_polygons_glyph_classes(handle), shape=(npoly * k,)).copy().reshape(npoly, k)
                                      if npoly > 0 else np.zeros((0, k), np.int32))
            L.ocrb_polygons_free(handle)
            self.image_offsets, self.point_offsets, self.xy, self.all_scores = io, po, xy, sc
        self._polygons = self._scores = None
        if not hasattr(self, "glyph_classes"):
            self.glyph_classes = None

    @property
    def num_images(self):
        return len(self.image_offsets) - 1

    @property
    def polygons(self):
        if self._polygons is None:
            io, po, xy = self.image_offsets, self.point_offsets, self.xy
            self._polygons = [[xy[po[p]:po[p + 1]] for p in range(io[b], io[b + 1])] for b in range(self.num_images)]
        return self._polygons

    @property
    def scores(self):
        if self._scores is None:
            io, sc = self.image_offsets, self.all_scores
            self._scores = [sc[io[b]:io[b + 1]] for b in range(self.num_images)]
        return self._scores

    def arrays(self):
        return self.image_offsets, self.point_offsets, self.xy, self.all_scores, self.stats

    @staticmethod
    def concat(parts):
        """Concatenates per-shard results in the given (= image index) order."""
        io, po, xy, sc, st = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)], [], [], []
        gc = [a.glyph_classes if isinstance(a, Polygons) else (a[5] if len(a) > 5 else None) for a in parts]
        for a in parts:
            pio, ppo, pxy, psc, pst = a.arrays() if isinstance(a, Polygons) else a[:5]
            io.append(pio[1:] + io[-1][-1])
            po.append(ppo[1:] + po[-1][-1])
            xy.append(pxy)
            sc.append(psc)
            st.append(pst)
        out = Polygons(arrays=(np.concatenate(io), np.concatenate(po), np.concatenate(xy) if xy else np.zeros((0, 2), np.uint32),
                               np.concatenate(sc) if sc else np.zeros(0), np.concatenate(st) if st else np.zeros((0, 5), np.int64)))
        if gc and all(g is not None for g in gc):
            out.glyph_classes = np.concatenate(gc)
        return out
