"""ORACLE-SIDE CHECKER (test infrastructure only): independent winding-number test of the Clipper
clean-up (see region_check.c).  Used by tests/test_union_region.py (oracle, CPU) and
tests/test_gpu_postproc.py (CUDA hook)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libregion_check.so")
_lib = None
S = 4  # samples per pixel and axis (first pass; see check_multires)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "region_check.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "libregion_check.so"])
        _lib = C.CDLL(_SO)
    return _lib


def winding_raster(pts, x0, y0, nx, ny):
    p = np.ascontiguousarray(np.asarray(pts, np.int32).reshape(-1, 2))
    out = np.empty((ny, nx), np.int16)
    lib().rc_winding_raster(p.ctypes.data_as(C.c_void_p), len(p), int(x0), int(y0), S, nx, ny, out.ctypes.data_as(C.c_void_p))
    return out


def _near_segments(x0, y0, nx, ny, segs, tol):
    """mask of samples within `tol` pixels of any of the segments [(ax, ay, bx, by)]."""
    zone = np.zeros((ny, nx), bool)
    for ax, ay, bx, by in segs:
        lo_i = max(0, int((min(ax, bx) - tol - x0) * S) - 1)
        hi_i = min(nx, int((max(ax, bx) + tol - x0) * S) + 2)
        lo_k = max(0, int((min(ay, by) - tol - y0) * S) - 1)
        hi_k = min(ny, int((max(ay, by) + tol - y0) * S) + 2)
        if lo_i >= hi_i or lo_k >= hi_k:
            continue
        xs = x0 + (2 * np.arange(lo_i, hi_i) + 1) / (2 * S)
        ys = y0 + (2 * np.arange(lo_k, hi_k) + 1) / (2 * S)
        X, Y = np.meshgrid(xs, ys)
        dx, dy = bx - ax, by - ay
        L = dx * dx + dy * dy
        t = np.clip(((X - ax) * dx + (Y - ay) * dy) / L, 0, 1) if L > 0 else np.zeros_like(X)
        d2 = (X - (ax + t * dx)) ** 2 + (Y - (ay + t * dy)) ** 2
        zone[lo_k:hi_k, lo_i:hi_i] |= d2 <= tol * tol
    return zone


def _on_raw_segment(raw, p, q):
    """does the emitted edge p-q lie on one segment of the raw path (then it is exact)?"""
    a = raw
    b = np.roll(raw, -1, 0)
    d = b - a
    on = np.ones(len(a), bool)
    for v in (p, q):
        w = v[None, :] - a
        on &= (d[:, 0] * w[:, 1] - d[:, 1] * w[:, 0]) == 0
        on &= (np.minimum(a, b) <= v[None, :]).all(1) & (v[None, :] <= np.maximum(a, b)).all(1)
    return bool(on.any())


def check(raw, emitted, tol=0.75, min_area_px=1.5):
    """raw: the rounded offset path; emitted: the polygon the implementation returned (or None).
    -> dict(ok, violations (samples), n_components, first_rule_ok, area_px)

    Rule: {winding(raw) > 0}, restricted to the connected component the emitted polygon covers
    and with its holes filled (the reference reads the exterior only, polygon.rs:35-40), must equal
    the inside of the emitted polygon at every sample farther than `tol` pixels from a FUZZY emitted
    edge.  An emitted edge is exact when it lies on a segment of the raw path; otherwise one of its
    end points is a crossing rounded to the integer grid (Clipper's IntersectPoint), the only place
    where a sliver narrower than the rounding distance may appear.  Differences thinner than 0.75 px
    (three samples) are ignored as well: they are below the resolution of the integer grid."""
    from scipy import ndimage

    raw = np.asarray(raw, np.int64).reshape(-1, 2)
    allp = raw if emitted is None else np.concatenate([raw, np.asarray(emitted, np.int64).reshape(-1, 2)])
    x0, y0 = int(allp[:, 0].min()) - 2, int(allp[:, 1].min()) - 2
    nx, ny = (int(allp[:, 0].max()) + 3 - x0) * S, (int(allp[:, 1].max()) + 3 - y0) * S
    R = winding_raster(raw, x0, y0, nx, ny) > 0
    if emitted is None:
        lab, n = ndimage.label(R)
        sizes = ndimage.sum(R, lab, np.arange(1, n + 1)) if n else np.zeros(0)
        big = [i + 1 for i in range(n) if sizes[i] >= min_area_px * S * S]
        return dict(ok=len(big) == 0, violations=int(sum(sizes[i - 1] for i in big)), n_components=len(big), first_rule_ok=True, area_px=0.0)
    em = np.asarray(emitted, np.int64).reshape(-1, 2)
    E = winding_raster(em, x0, y0, nx, ny) != 0
    segs = []
    for i in range(len(em)):
        j = (i + 1) % len(em)
        if not _on_raw_segment(raw, em[i], em[j]):
            segs.append((float(em[i, 0]), float(em[i, 1]), float(em[j, 0]), float(em[j, 1])))
    zone = None

    def outside_zone(diff):
        nonlocal zone
        if not diff.any():
            return diff
        if zone is None:
            zone = _near_segments(x0, y0, nx, ny, segs, tol)
        diff = diff & ~zone
        # features thinner than 3 samples (0.75 px) cannot survive the integer grid either way: the raw
        # path itself is only defined up to its rounding.  What remains after an opening is a real
        # difference of regions.
        k = max(3, int(round(0.75 * S)))
        return ndimage.binary_opening(diff, structure=np.ones((k, k), bool)) if diff.any() else diff

    area = E.sum() / (S * S)
    # fast path: one component, no holes
    if not outside_zone(R ^ E).any():
        return dict(ok=True, violations=0, n_components=1, first_rule_ok=True, area_px=area)
    # several components and / or holes.  Regions that touch in a single point may come out of
    # Clipper's sweep as one self-touching ring or as separate rings; the emitted polygon may
    # therefore cover one component or several, but each one entirely or not at all.
    lab, n = ndimage.label(R)  # 4-connectivity
    if n == 0:
        return dict(ok=False, violations=int(E.sum()), n_components=0, first_rule_ok=True, area_px=area)
    idx = np.arange(1, n + 1)
    sizes = ndimage.sum(R, lab, idx)
    inside = ndimage.sum(E, lab, idx)
    covered = [i for i in idx if inside[i - 1] * 2 > sizes[i - 1]]
    comp = ndimage.binary_fill_holes(np.isin(lab, covered))
    v = int(outside_zone(comp ^ E).sum())
    big = [i for i in idx if sizes[i - 1] >= min_area_px * S * S]
    # "first polygon" statistic: does the emitted polygon own the largest-y sample among the components
    # that are not slivers?
    first_ok = True
    if big:
        lowest = max(big, key=lambda i: (np.nonzero((lab == i).any(1))[0].max(), -np.nonzero((lab == i).any(0))[0].min()))
        first_ok = lowest in covered
    return dict(ok=v == 0, violations=v, n_components=len(big), first_rule_ok=first_ok, area_px=area)


def check_multires(raw, emitted, levels=(4, 16, 32)):
    """check() at increasing sampling densities: wedge-shaped channels and necks narrower than a sample
    make the SAMPLED region connected where the exact one is not (or the reverse); a finer grid resolves
    them.  Passes at the first density that passes; the result carries the density used."""
    global S
    r = None
    try:
        for lv in levels:
            S = lv
            r = check(raw, emitted)
            r["S"] = lv
            if r["ok"]:
                break
    finally:
        S = 4
    return r


def random_dp_polygon(rng, kind=None):
    """Random polygons of the kinds Douglas-Peucker leaves behind (and worse): convex, concave stars,
    spikes, near-collinear runs, thin slivers, self-intersecting."""
    kind = kind if kind is not None else rng.integers(0, 7)
    n = int(rng.integers(4, 16))
    cx, cy = rng.integers(200, 600, 2)
    if kind == 0:  # rotated rectangle with jitter
        w, h, a = rng.uniform(8, 150), rng.uniform(5, 60), rng.uniform(0, np.pi)
        base = np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) / 2
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        pts = base @ R.T + rng.uniform(-2, 2, (4, 2))
    elif kind == 1:  # star-shaped, mildly concave
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        r = rng.uniform(15, 80) * rng.uniform(0.5, 1.0, n)
        pts = np.stack([r * np.cos(ang), r * np.sin(ang)], 1)
    elif kind == 2:  # spiky star
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        r = rng.uniform(20, 90) * np.where(np.arange(n) % 2 == 0, 1.0, rng.uniform(0.1, 0.5, n))
        pts = np.stack([r * np.cos(ang), r * np.sin(ang)], 1)
    elif kind == 3:  # near-collinear runs on a long thin shape
        L, h = rng.uniform(40, 200), rng.uniform(3, 20)
        k = n // 2 + 1
        top = np.stack([np.sort(rng.uniform(0, L, k)), rng.uniform(-0.7, 0.7, k)], 1)
        bot = np.stack([np.sort(rng.uniform(0, L, k))[::-1], h + rng.uniform(-0.7, 0.7, k)], 1)
        a = rng.uniform(0, np.pi)
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        pts = np.concatenate([top, bot]) @ R.T
    elif kind == 4:  # L / U shapes: narrow notches that an expansion closes
        w, h, t = rng.uniform(30, 120), rng.uniform(30, 120), rng.uniform(2, 12)
        g = rng.uniform(2, 25)
        pts = np.array([[0, 0], [w, 0], [w, h], [w / 2 + g / 2, h], [w / 2 + g / 2, t], [w / 2 - g / 2, t], [w / 2 - g / 2, h], [0, h]], float)
        a = rng.uniform(0, np.pi)
        R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
        pts = pts @ R.T
    elif kind == 5:  # star with a few neighbouring vertices swapped: the mild self-intersections DP can leave
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        r = rng.uniform(20, 80) * rng.uniform(0.5, 1.0, n)
        pts = np.stack([r * np.cos(ang), r * np.sin(ang)], 1)
        for _ in range(int(rng.integers(1, 3))):
            i = int(rng.integers(0, n))
            pts[[i, (i + 1) % n]] = pts[[(i + 1) % n, i]]
    else:  # unsorted angles: chaotic self-intersections with sub-pixel slivers (statistics only)
        ang = rng.uniform(0, 2 * np.pi, n)
        r = rng.uniform(15, 70) * rng.uniform(0.4, 1.0, n)
        pts = np.stack([r * np.cos(ang), r * np.sin(ang)], 1)
    pts = np.round(pts + (cx, cy)).astype(np.int32)
    if rng.random() < 0.5:
        pts = pts[::-1].copy()
    return pts
