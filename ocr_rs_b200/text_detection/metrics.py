"""text_detection/metrics.rs mirror (inference part, metrics.rs:32-184)."""
import ctypes as C

import numpy as np

from .. import _ffi


def _ctx(ctx):
    return ctx if ctx is not None else _ffi.default_context()


class PolygonScores:
    """metrics.rs:32-35: polygons[b] = list of uint32 [m,2] arrays, scores[b] = f64 array."""

    def __init__(self, polys: "_ffi.Polygons"):
        self.polygons = polys.polygons
        self.scores = polys.scores
        self.stats = polys.stats


def _params(thresh=None, box_thresh=None, min_size=None, unclip_factor=None):
    p = _ffi.PostprocParams()
    _ffi.lib().ocrb_postproc_default_params(C.byref(p))
    if thresh is not None:
        p.thresh = thresh
    if box_thresh is not None:
        p.box_thresh = box_thresh
    if min_size is not None:
        p.min_size = min_size
    if unclip_factor is not None:
        p.unclip_factor = unclip_factor
    return p


def binarize(pred, thresh, ctx=None):
    """metrics.rs:129-131: pred.gt(thresh).to_kind(Uint8)."""
    ctx = _ctx(ctx)
    if hasattr(pred, "data_ptr"):
        import torch
        pred = pred.contiguous()
        out = torch.empty(pred.shape, dtype=torch.uint8, device=pred.device)
        n = pred.numel()
    else:
        pred = np.ascontiguousarray(pred, np.float32)
        out = np.empty(pred.shape, np.uint8)
        n = pred.size
    _ffi.check(_ffi.lib().ocrb_binarize(ctx.handle, _ffi.ptr(pred), n, float(thresh), _ffi.ptr(out)))
    return out


def box_score_fast(bitmap, points, ctx=None):
    """metrics.rs:150-184: bitmap float32 [H,W] probability map, points [(x,y)] -> f64 mean."""
    ctx = _ctx(ctx)
    pts = np.ascontiguousarray(np.asarray(points, np.int32).reshape(-1, 2))
    s = C.c_double()
    if not hasattr(bitmap, "data_ptr"):
        bitmap = np.ascontiguousarray(bitmap, np.float32)
    _ffi.check(_ffi.lib().ocrb_box_score_fast(ctx.handle, _ffi.ptr(bitmap), int(bitmap.shape[-2]), int(bitmap.shape[-1]),
                                              _ffi.ptr(pts), len(pts), C.byref(s)))
    return s.value


def get_min_area_bounding_box(contour, ctx=None):
    """metrics.rs:133-148 -> (box int32 [4,2], short side f64)."""
    ctx = _ctx(ctx)
    pts = np.ascontiguousarray(np.asarray(contour, np.int32).reshape(-1, 2))
    box = np.empty((4, 2), np.int32)
    s = C.c_double()
    _ffi.check(_ffi.lib().ocrb_min_area_bounding_box(ctx.handle, _ffi.ptr(pts), len(pts), _ffi.ptr(box), C.byref(s)))
    return box, s.value


def get_polygons_from_bitmap(pred, bitmap, adjust_values, ctx=None, **kw):
    """metrics.rs:58-127: pred float32 [1,H,W] (or [H,W]), bitmap uint8 same shape,
    adjust_values [2] -> (list of uint32 [m,2] polygons, f64 scores)."""
    ctx = _ctx(ctx)
    H, W = int(pred.shape[-2]), int(pred.shape[-1])
    if not hasattr(pred, "data_ptr"):
        pred = np.ascontiguousarray(pred, np.float32)
    if not hasattr(bitmap, "data_ptr"):
        bitmap = np.ascontiguousarray(bitmap, np.uint8)
    adj = np.ascontiguousarray(np.asarray(adjust_values, np.float64).reshape(2))
    h = _ffi.c_p()
    p = _params(**kw)
    _ffi.check(_ffi.lib().ocrb_get_polygons_from_bitmap(ctx.handle, _ffi.ptr(pred), _ffi.ptr(bitmap), _ffi.ptr(adj), H, W,
                                                        C.byref(p), C.byref(h)))
    res = _ffi.Polygons(h)
    return res.polygons[0], res.scores[0]


def get_boxes_and_box_scores(pred, adjust_values, ctx=None, return_raw=False, **kw):
    """metrics.rs:37-56: pred float32 [B,1,H,W], adjust_values [B,2] f64 -> PolygonScores."""
    ctx = _ctx(ctx)
    B, H, W = int(pred.shape[0]), int(pred.shape[-2]), int(pred.shape[-1])
    if not hasattr(pred, "data_ptr"):
        pred = np.ascontiguousarray(pred, np.float32)
    adj = np.ascontiguousarray(np.asarray(adjust_values, np.float64).reshape(B, 2))
    h = _ffi.c_p()
    p = _params(**kw)
    _ffi.check(_ffi.lib().ocrb_get_boxes_and_box_scores(ctx.handle, _ffi.ptr(pred), _ffi.ptr(adj), B, H, W, C.byref(p),
                                                        C.byref(h)))
    raw = _ffi.Polygons(h)
    return raw if return_raw else PolygonScores(raw)


# ---- contour-stage hooks (imageproc find_contours / approximate_polygon_dp) ---------------
def ccl_labels(bitmap, ctx=None):
    """8-connected foreground labels in raster order (== scipy.ndimage.label(ones(3,3)))."""
    ctx = _ctx(ctx)
    bm = np.ascontiguousarray(bitmap, np.uint8)
    b3 = bm.reshape((-1,) + bm.shape[-2:])
    out = np.empty(b3.shape, np.int32)
    n = np.zeros(b3.shape[0], np.int32)
    _ffi.check(_ffi.lib().ocrb_ccl_labels(ctx.handle, _ffi.ptr(b3), b3.shape[0], b3.shape[1], b3.shape[2], _ffi.ptr(out), _ffi.ptr(n)))
    return out.reshape(bm.shape), (n if bm.ndim == 3 else int(n[0]))


def find_contours(bitmap, ctx=None):
    """-> (list of int32 [n_i,2] chains, types uint8[n]) in the reference's order."""
    ctx = _ctx(ctx)
    bm = np.ascontiguousarray(bitmap, np.uint8)
    H, W = bm.shape
    nc, npts = _ffi.i64(0), _ffi.i64(0)
    L = _ffi.lib()
    _ffi.check(L.ocrb_find_contours(ctx.handle, _ffi.ptr(bm), H, W, None, None, 0, None, 0, C.byref(nc), C.byref(npts)))
    offs = np.zeros(nc.value + 1, np.int64)
    types = np.zeros(max(nc.value, 1), np.uint8)
    xy = np.zeros((max(npts.value, 1), 2), np.int32)
    _ffi.check(L.ocrb_find_contours(ctx.handle, _ffi.ptr(bm), H, W, _ffi.ptr(offs), _ffi.ptr(types), nc.value,
                                    _ffi.ptr(xy), npts.value, C.byref(nc), C.byref(npts)))
    return [xy[offs[i]:offs[i + 1]].copy() for i in range(nc.value)], types[: nc.value]


def approx_polygon(chain, ctx=None):
    """metrics.rs:87-95 on one chain."""
    ctx = _ctx(ctx)
    c = np.ascontiguousarray(np.asarray(chain, np.int32).reshape(-1, 2))
    out = np.empty((len(c) + 2, 2), np.int32)
    n = _ffi.i64(0)
    _ffi.check(_ffi.lib().ocrb_approx_polygon(ctx.handle, _ffi.ptr(c), len(c), _ffi.ptr(out), len(out), C.byref(n)))
    return out[: n.value].copy()


# ---- evaluation metrics (metrics.rs:191-394; host code inside libocrb, no device needed) -----------------
class MetricsItem:
    """metrics.rs:22-30"""
    __slots__ = ("precision", "recall", "hmean", "gt_care", "det_care", "det_matched")

    def __init__(self, precision, recall, hmean, gt_care, det_care, det_matched):
        self.precision, self.recall, self.hmean = precision, recall, hmean
        self.gt_care, self.det_care, self.det_matched = gt_care, det_care, det_matched

    def __repr__(self):
        return (f"MetricsItem(precision={self.precision}, recall={self.recall}, hmean={self.hmean}, gt_care={self.gt_care}, "
                f"det_care={self.det_care}, det_matched={self.det_matched})")


def _csr(polys):
    offs = np.zeros(len(polys) + 1, np.int64)
    for i, p in enumerate(polys):
        offs[i + 1] = offs[i] + len(p)
    xy = (np.concatenate([np.asarray(p, np.uint32).reshape(-1, 2) for p in polys]) if len(polys) else np.zeros((0, 2), np.uint32))
    return offs, np.ascontiguousarray(xy, np.uint32)


def polygon_iou(a, b):
    """get_intersection_over_union (metrics.rs:387-389) -> (intersection area, IoU)"""
    a = np.ascontiguousarray(np.asarray(a, np.uint32).reshape(-1, 2))
    b = np.ascontiguousarray(np.asarray(b, np.uint32).reshape(-1, 2))
    inter, iou = C.c_double(), C.c_double()
    _ffi.check(_ffi.lib().ocrb_polygon_iou(_ffi.ptr(a), len(a), _ffi.ptr(b), len(b), C.byref(inter), C.byref(iou)))
    return inter.value, iou.value


def evaluate_image(gt_points, ignore_flags, pred):
    """metrics.rs:251-372: gt_points / pred = lists of polygons [(x, y)], ignore_flags = list of bool -> MetricsItem"""
    go, gxy = _csr(gt_points)
    do, dxy = _csr(pred)
    ig = np.ascontiguousarray(np.asarray(ignore_flags, bool).astype(np.uint8))
    item = _ffi.MetricsItem()
    _ffi.check(_ffi.lib().ocrb_evaluate_image(_ffi.ptr(go), _ffi.ptr(gxy), len(gt_points), _ffi.ptr(ig) if len(ig) else None,
                                              _ffi.ptr(do), _ffi.ptr(dxy), len(pred), C.byref(item)))
    return MetricsItem(item.precision, item.recall, item.hmean, item.gt_care, item.det_care, item.det_matched)


def validate_measure(polygons, ignore_tags, pred, scores):
    """metrics.rs:191-218: per image, the predictions with score >= 0.6 against the ground truth"""
    box_thresh = 0.6
    out = []
    for gt, ig, pr, sc in zip(polygons, ignore_tags, pred, scores):
        kept = [p for s, p in zip(sc, pr) if s >= box_thresh]
        out.append(evaluate_image(gt, ig, kept))
    return out


def combine_results(results):
    """metrics.rs:229-249 -> (precision, recall, hmean)"""
    arr = (_ffi.MetricsItem * len(results))(*[_ffi.MetricsItem(r.precision, r.recall, r.hmean, r.gt_care, r.det_care, r.det_matched) for r in results])
    p, r, h = C.c_double(), C.c_double(), C.c_double()
    _ffi.check(_ffi.lib().ocrb_combine_results(arr, len(results), C.byref(p), C.byref(r), C.byref(h)))
    return p.value, r.value, h.value


def gather_measure(metrics):
    """metrics.rs:220-227: metrics = list (batches) of lists of MetricsItem"""
    return combine_results([m for batch in metrics for m in batch])
