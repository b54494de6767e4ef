"""End-to-end polygon comparison against the oracle's own end-to-end path (torch-CPU map -> C post-processing),
north_star: "end-to-end polygons identical at IoU >= 0.99".

A polygon can only be required to match where the decision that created it is stable under the map tolerance of
the arithmetic mode.  The exclusion rule, stated once:

  map tolerance  tol = 1e-2 (BF16 mode) / 1e-4 (FP32 mode) on the probability of a gain-1 head, i.e. a logit
  tolerance of 4 * tol (d logit = dp / (p (1 - p)) >= 4 dp); the structured test head multiplies the logit by
  `gain` (64), so the device logit may differ from the oracle's by  L = 4 * tol * gain.

  A polygon is UNDECIDABLE when, inside its bounding box grown by 3 px (map coordinates), some pixel of the oracle's
  map has |logit(p) - logit(0.6)| <= L (its bitmap bit may flip: binarize threshold, metrics.rs:38), or when its
  box score is within sigmoid'(.)-scaled reach of the 0.7 filter (metrics.rs:100): |score - 0.7| <= the largest
  probability change a logit shift of L can cause, or when its min-area-rect short side is within 1 px of the
  size filter (metrics.rs:105).

Every other polygon is DECIDABLE and must have a partner at IoU >= 0.99 — in both directions."""
import numpy as np

from oracle import postproc as pp


def _logit(p):
    p = np.clip(p.astype(np.float64), 1e-300, 1 - 1e-16)
    return np.log(p / (1 - p))


def _undecidable(ref_map, poly, score, adjust, L):
    H, W = ref_map.shape
    pts = np.asarray(poly, np.float64) * np.asarray(adjust, np.float64)[None, :]
    x0, y0 = np.floor(pts.min(0)).astype(int) - 3
    x1, y1 = np.ceil(pts.max(0)).astype(int) + 4
    win = ref_map[max(0, y0):min(H, y1), max(0, x0):min(W, x1)]
    if win.size and (np.abs(_logit(win) - _logit(np.array(0.6))) <= L).any():
        return True
    if score is not None and np.isfinite(score):
        # largest probability move under a logit shift of L, at this score
        lo = 1 / (1 + np.exp(-(_logit(np.array(score)) - L)))
        hi = 1 / (1 + np.exp(-(_logit(np.array(score)) + L)))
        if lo <= 0.7 <= hi:
            return True
    _, sside = pp.min_area_bounding_box(np.round(pts).astype(np.int32))
    return abs(sside - 5.0) <= 1.0


def compare(ref_map, got_polys, got_scores, adjust, tol, gain):
    """-> dict(matched, undecidable, failures=[...]) for one image."""
    L = 4.0 * tol * gain
    exp_p, exp_s = pp.polygons_from_bitmap(ref_map, pp.binarize(ref_map, 0.6), tuple(adjust))
    out = dict(expected=len(exp_p), got=len(got_polys), matched=0, undecidable=0, failures=[])
    used = set()
    for e, s in zip(exp_p, exp_s):
        best, bi = 0.0, -1
        for i, a in enumerate(got_polys):
            iou = pp.polygon_iou(e, a)
            if iou > best:
                best, bi = iou, i
        if best >= 0.99:
            out["matched"] += 1
            used.add(bi)
        elif _undecidable(ref_map, e, s, adjust, L):
            out["undecidable"] += 1
        else:
            out["failures"].append(("missing", e.tolist(), float(s), best))
    for i, a in enumerate(got_polys):
        if i in used:
            continue
        if max((pp.polygon_iou(a, e) for e in exp_p), default=0.0) >= 0.99:
            continue
        if _undecidable(ref_map, a, None if got_scores is None else got_scores[i], adjust, L):
            out["undecidable"] += 1
        else:
            out["failures"].append(("extra", np.asarray(a).tolist(), None if got_scores is None else float(got_scores[i]), 0.0))
    return out
