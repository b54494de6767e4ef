"""ORACLE (test infrastructure only): ctypes front-end of oracle/postproc_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.  See the header of postproc_oracle.c for what is restated and how it is
pinned (reference goldens metrics.rs:406-646).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpostproc_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "postproc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libpostproc_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_arc_length.restype = C.c_double
        _lib.orc_box_score.restype = C.c_double
        _lib.orc_min_area_bounding_box.restype = C.c_double
        _lib.orc_approx_dp.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _pts(points):
    a = np.ascontiguousarray(np.asarray(points, np.int32).reshape(-1, 2))
    return a


def binarize(pred, thresh=0.6):
    pred = np.ascontiguousarray(pred, np.float32)
    out = np.empty(pred.shape, np.uint8)
    lib().orc_binarize(_p(pred, C.c_float), C.c_int64(pred.size), C.c_double(thresh), _p(out, C.c_uint8))
    return out


def find_contours(bitmap):
    """-> (list of [n_i,2] int32 arrays (x,y), types uint8[n] 0=outer 1=hole)."""
    bm = np.ascontiguousarray(bitmap, np.uint8)
    H, W = bm.shape
    cap_pts = W * H * 2 + 16
    cap_c = W * H // 2 + 16
    pts = np.empty((cap_pts, 2), np.int32)
    offs = np.empty(cap_c + 1, np.int64)
    types = np.empty(cap_c, np.uint8)
    n = lib().orc_find_contours(_p(bm, C.c_uint8), W, H, _p(pts, C.c_int32), C.c_int64(cap_pts),
                                _p(offs, C.c_int64), _p(types, C.c_uint8), cap_c)
    if n < 0:
        raise RuntimeError("oracle find_contours capacity")
    return [pts[offs[i]:offs[i + 1]].copy() for i in range(n)], types[:n].copy()


def arc_length(chain, closed=True):
    c = _pts(chain)
    return float(lib().orc_arc_length(_p(c, C.c_int32), C.c_int64(len(c)), int(closed)))


def approx_dp(chain, eps, closed=True):
    c = _pts(chain)
    out = np.empty((len(c) + 2, 2), np.int32)
    m = lib().orc_approx_dp(_p(c, C.c_int32), C.c_int64(len(c)), C.c_double(eps), int(closed), _p(out, C.c_int32))
    return out[:m].copy()


def dp_polygon(chain):
    """metrics.rs:87-95: eps = 1% arc length (0 -> 0.01), DP, drop duplicated last point."""
    eps = 0.01 * arc_length(chain, True)
    if eps == 0.0:
        eps = 0.01
    p = approx_dp(chain, eps, True)
    if len(p) > 1 and (p[0] == p[-1]).all():
        p = p[:-1]
    return p


def box_score(pred, points, return_count=False):
    pred = np.ascontiguousarray(pred, np.float32)
    p = _pts(points)
    cnt = C.c_int64(0)
    s = lib().orc_box_score(_p(pred, C.c_float), pred.shape[-2], pred.shape[-1], _p(p, C.c_int32), len(p), C.byref(cnt))
    return (float(s), cnt.value) if return_count else float(s)


def offset_raw(points, delta):
    p = _pts(points)
    out = np.empty((3 * len(p) + 3, 2), np.int32)
    m = lib().orc_offset_raw(_p(p, C.c_int32), len(p), C.c_double(delta), _p(out, C.c_int32))
    return out[:m].copy()


def clip_polygon(points, factor, shrink, return_distance=False):
    """polygon.rs:13-42 -> [m,2] int32 or None."""
    p = _pts(points)
    cap = 6 * len(p) + 32
    out = np.empty((cap, 2), np.int32)
    d = C.c_double(0)
    m = lib().orc_clip_polygon(_p(p, C.c_int32), len(p), C.c_double(factor), int(bool(shrink)), _p(out, C.c_int32), cap, C.byref(d))
    res = out[:m].copy() if m > 0 else None
    return (res, d.value) if return_distance else res


def expand_polygon(points, factor=2.0, return_distance=False):
    """polygon.rs:51-56 -> [m,2] int32 or None."""
    return clip_polygon(points, factor, False, return_distance)


def shrink_polygon(points, factor):
    """polygon.rs:44-49 (training-side caller image_ops.rs:265; used here to pin the Clipper restatement)."""
    return clip_polygon(points, factor, True)


def union_of_path(raw):
    """Clipper's clean-up of one closed (offset) path -> [m,2] int32 or None."""
    p = _pts(raw)
    cap = 4 * len(p) + 16
    out = np.empty((cap, 2), np.int32)
    m = lib().orc_union_of_path(_p(p, C.c_int32), len(p), _p(out, C.c_int32), cap)
    return out[:m].copy() if m > 0 else None


def draw_polygon(canvas, points, value):
    """imageproc draw_polygon_mut on a [H,W] u8 canvas, in place."""
    assert canvas.dtype == np.uint8 and canvas.flags.c_contiguous
    p = _pts(points)
    lib().orc_draw_polygon(_p(canvas, C.c_uint8), canvas.shape[1], canvas.shape[0], _p(p, C.c_int32), len(p), int(value))
    return canvas


MIN_TEXT_SIZE = 8  # image_ops.rs:42


def generate_gt_and_mask_images(polygons, adjust_x, adjust_y, target_dim):
    """image_ops.rs:222-277 (dataset preparation; not on the hot path): the ground-truth map is every
    polygon shrunk by 1 - 0.5^2 and filled, the mask blanks ignored polygons.  -> (gt, mask, ignore_flags)"""
    width, height = target_dim
    gt = np.zeros((height, width), np.uint8)
    mask = np.full((height, width), 255, np.uint8)
    flags = []
    for poly in polygons:
        pts = np.asarray(poly, np.int64)
        pw, ph = pts[:, 0].max() - pts[:, 0].min(), pts[:, 1].max() - pts[:, 1].min()
        vals = np.stack([(pts[:, 0].astype(np.float64) * adjust_x).astype(np.int32),
                         (pts[:, 1].astype(np.float64) * adjust_y).astype(np.int32)], 1)
        if len(vals) < 4:
            flags.append(True)
            continue
        if min(pw, ph) < MIN_TEXT_SIZE:
            draw_polygon(mask, vals, 0)
            flags.append(True)
            continue
        sh = shrink_polygon(vals, 1.0 - 0.5 ** 2)
        if sh is None:
            draw_polygon(mask, vals, 0)
            flags.append(True)
        else:
            draw_polygon(gt, sh, 255)
            flags.append(False)
    return gt, mask, flags


def min_area_bounding_box(points):
    """metrics.rs:133-148 -> (box [4,2] int32, short side f64)."""
    p = _pts(points)
    box = np.empty((4, 2), np.int32)
    s = lib().orc_min_area_bounding_box(_p(p, C.c_int32), len(p), _p(box, C.c_int32))
    return box, float(s)


def polygons_from_bitmap(pred, bitmap, adjust=(1.0, 1.0), box_thresh=0.7, min_size=5.0,
                         unclip=2.0, return_stats=False, return_boxes=False):
    """metrics.rs:58-127 for one image -> (list of [m,2] uint32 arrays, scores f64[n])."""
    pred = np.ascontiguousarray(pred, np.float32)
    bm = np.ascontiguousarray(bitmap, np.uint8)
    H, W = bm.shape
    max_polys = W * H // 2 + 16
    max_pts = 4 * W * H + 64
    offs = np.empty(max_polys + 1, np.int64)
    xy = np.empty((max_pts, 2), np.uint32)
    scores = np.empty(max_polys, np.float64)
    stats = np.zeros(5, np.int64)
    boxes = np.empty((max_polys if return_boxes else 0, 4, 2), np.int32)
    n = lib().orc_polygons_from_bitmap(
        _p(pred, C.c_float), _p(bm, C.c_uint8), H, W, C.c_double(adjust[0]), C.c_double(adjust[1]),
        C.c_double(box_thresh), C.c_double(min_size), C.c_double(unclip), max_polys, C.c_int64(max_pts),
        _p(offs, C.c_int64), _p(xy, C.c_uint32), _p(scores, C.c_double), _p(stats, C.c_int64),
        _p(boxes, C.c_int32) if return_boxes else None)
    if n < 0:
        raise RuntimeError("oracle polygons_from_bitmap capacity")
    polys = [xy[offs[i]:offs[i + 1]].copy() for i in range(n)]
    if return_boxes:
        return polys, scores[:n].copy(), boxes[:n].copy()
    if return_stats:
        return polys, scores[:n].copy(), stats
    return polys, scores[:n].copy()


def boxes_and_box_scores(pred, adjust_values, thresh=0.6):
    """metrics.rs:37-56: pred [B,1,H,W] f32, adjust [B,2] -> (polygons per image, scores per image)."""
    pred = np.asarray(pred, np.float32)
    seg = binarize(pred, thresh)
    polys, scores = [], []
    for b in range(pred.shape[0]):
        p, s = polygons_from_bitmap(pred[b, 0], seg[b, 0], tuple(np.asarray(adjust_values, np.float64)[b]))
        polys.append(p)
        scores.append(s)
    return polys, scores


def preprocess(rgba, W, H, norm_first=False):
    """image_ops.rs:188-220 minus the file decode: RGBA8 [h,w,4] -> (u8 [H,W], adjust_x, adjust_y)."""
    rgba = np.ascontiguousarray(rgba, np.uint8)
    sh, sw = rgba.shape[:2]
    dst = np.empty((H, W), np.uint8)
    rw, rh = C.c_int(0), C.c_int(0)
    lib().orc_set_resize_variant(int(norm_first))
    lib().orc_preprocess(_p(rgba, C.c_uint8), sw, sh, W, H, _p(dst, C.c_uint8), C.byref(rw), C.byref(rh))
    return dst, rw.value / sw, rh.value / sh


def polygon_iou(a, b):
    """IoU of two simple polygons by even-odd rasterisation on a 4x supersampled grid
    (host-side helper for the north_star 'IoU >= 0.99' check; exact clipping is
    metrics.rs:382-394 territory and out of the kernel path)."""
    import cv2
    a = np.asarray(a, np.int64)
    b = np.asarray(b, np.int64)
    lo = np.minimum(a.min(0), b.min(0)) - 1
    hi = np.maximum(a.max(0), b.max(0)) + 2
    S = 4
    size = ((hi - lo) * S).astype(int)
    ma = np.zeros((size[1], size[0]), np.uint8)
    mb = np.zeros_like(ma)
    cv2.fillPoly(ma, [((a - lo) * S).astype(np.int32)], 1)
    cv2.fillPoly(mb, [((b - lo) * S).astype(np.int32)], 1)
    inter = np.logical_and(ma, mb).sum()
    union = np.logical_or(ma, mb).sum()
    return inter / union if union else 1.0


def crop_glyphs(image, box, k):
    """crop spec v1 (see orc_crop_glyphs): u8 [H,W] image, box [4,2] (TL,TR,BR,BL) -> u8 [k,784]."""
    img = np.ascontiguousarray(image, np.uint8)
    b = _pts(box)
    out = np.empty((k, 784), np.uint8)
    lib().orc_crop_glyphs(_p(img, C.c_uint8), img.shape[1], img.shape[0], _p(b, C.c_int32), int(k), _p(out, C.c_uint8))
    return out


def kept_boxes_from_bitmap(pred, bitmap, box_thresh=0.7, min_size=5.0, unclip=2.0):
    """The min-area boxes (map coordinates, TL,TR,BR,BL) of the polygons get_polygons_from_bitmap keeps, in its
    order: contours -> DP -> score filter -> unclip -> size filter (metrics.rs:78-108)."""
    pred = np.ascontiguousarray(pred, np.float32)
    cs, _ = find_contours(bitmap)
    boxes = []
    for c in cs:
        dp = dp_polygon(c)
        if len(dp) < 4:
            continue
        if box_thresh > box_score(pred, dp):
            continue
        ex = expand_polygon(dp, unclip)
        if ex is None:
            continue
        box, sside = min_area_bounding_box(ex)
        if sside < min_size:
            continue
        boxes.append(box)
    return boxes
