"""CPU: evaluation metrics (metrics.rs:191-394) — the reference's own known-answer tests (metrics.rs:648-901), and the
polygon intersection they rest on against an independent raster estimate.  Host code inside libocrb: no device needed."""
import numpy as np

from ocr_rs_b200.text_detection import metrics as M

EPS = np.finfo(np.float64).eps
GT = [[(0, 0), (10, 0), (10, 10), (0, 10)], [(20, 20), (30, 20), (30, 30), (20, 30)]]
PRED1 = [(1, 1), (10, 0), (10, 10), (0, 10)]


def test_evaluate_image_one_matching_polygon():  # metrics.rs:648-678
    m = M.evaluate_image(GT, [False, False], [PRED1])
    assert (m.gt_care, m.det_care, m.det_matched) == (2, 1, 1)
    assert abs(m.precision - 1.) < EPS and abs(m.recall - 0.5) < EPS and abs(m.hmean - 0.6666666666666666) < EPS


def test_evaluate_image_with_ignored_polygons():  # metrics.rs:680-710
    m = M.evaluate_image(GT, [True, True], [PRED1])
    assert (m.gt_care, m.det_care, m.det_matched) == (0, 0, 0)
    assert abs(m.precision - 1.) < EPS and abs(m.recall - 1.) < EPS and abs(m.hmean - 1.) < EPS


def test_evaluate_image_with_both_matched_polygons():  # metrics.rs:712-748
    m = M.evaluate_image(GT, [False, False], [PRED1, GT[1]])
    assert (m.gt_care, m.det_care, m.det_matched) == (2, 2, 2)
    assert abs(m.precision - 1.) < EPS and abs(m.recall - 1.) < EPS and abs(m.hmean - 1.) < EPS


def test_validate_measure():  # metrics.rs:750-812
    pred = [[PRED1], [[(45, 61), (47, 41), (60, 60), (39, 48)]]]
    ms = M.validate_measure([GT, GT], [[False, False], [False, False]], pred, [[0.9], [0.9]])
    assert len(ms) == 2
    assert (ms[0].gt_care, ms[0].det_care, ms[0].det_matched) == (2, 1, 1)
    assert abs(ms[0].precision - 1.) < EPS and abs(ms[0].recall - 0.5) < EPS and abs(ms[0].hmean - 0.6666666666666666) < EPS
    assert (ms[1].gt_care, ms[1].det_care, ms[1].det_matched) == (2, 1, 0)
    assert ms[1].precision < EPS and ms[1].recall < EPS and ms[1].hmean < EPS
    # the 0.6 score filter (metrics.rs:197, :207-211)
    low = M.validate_measure([GT], [[False, False]], [[PRED1]], [[0.59]])[0]
    assert (low.det_care, low.det_matched) == (0, 0)


ITEMS = [M.MetricsItem(1., 0.5, 0.6666666666666666, 2, 1, 1), M.MetricsItem(1., 1., 1., 0, 0, 0),
         M.MetricsItem(1., 1., 1., 2, 2, 2), M.MetricsItem(0.3333333333333333, 0.2, 0.25, 5, 3, 1)]


def test_combine_results():  # metrics.rs:814-856 (assert_eq! on the tuple)
    assert M.combine_results(ITEMS) == (0.6666666666666666, 0.4444444444444444, 0.5333333333333333)
    assert M.combine_results([]) == (0., 0., 0.)


def test_gather_measure():  # metrics.rs:858-901
    assert M.gather_measure([ITEMS[:2], ITEMS[2:]]) == (0.6666666666666666, 0.4444444444444444, 0.5333333333333333)


def test_polygon_iou_exact_cases():
    sq = [(0, 0), (10, 0), (10, 10), (0, 10)]
    assert M.polygon_iou(sq, sq) == (100.0, 1.0)
    assert M.polygon_iou(sq, [(5, 0), (15, 0), (15, 10), (5, 10)]) == (50.0, 50.0 / 150.0)
    assert M.polygon_iou(sq, [(20, 20), (30, 20), (30, 30), (20, 30)]) == (0.0, 0.0)
    assert M.polygon_iou(sq, sq[::-1])[0] == 100.0  # orientation does not matter
    # concave: an L against the square that fills its notch
    L = [(0, 0), (10, 0), (10, 4), (4, 4), (4, 10), (0, 10)]
    assert M.polygon_iou(L, [(4, 4), (10, 4), (10, 10), (4, 10)])[0] == 0.0
    assert M.polygon_iou(L, sq) == (64.0, 0.64)
    # the KAT pair behind metrics.rs:648-678: pred misses the triangle (0,0),(10,0),(1,1) ... of the square
    inter, iou = M.polygon_iou(sq, PRED1)
    assert abs(inter - 90.0) < 1e-9 and abs(iou - 0.9) < 1e-12


def test_polygon_iou_against_raster_estimate():
    """random star-shaped polygons (thin spikes included): exact areas vs a 16x supersampled winding-number raster
    (oracle/region_check.c — scanline accumulation, shares nothing with the slab decomposition in csrc/eval.cu)"""
    from oracle import region_check as rc
    rng = np.random.default_rng(0)
    rc.S = 16
    try:
        for _ in range(200):
            polys = []
            for _k in range(2):
                n = int(rng.integers(3, 12))
                ang = np.sort(rng.uniform(0, 2 * np.pi, n))
                r = rng.uniform(10, 60, n)
                c = rng.uniform(80, 120, 2)
                polys.append(np.stack([c[0] + r * np.cos(ang), c[1] + r * np.sin(ang)], 1).round().astype(np.uint32))
            inter, iou = M.polygon_iou(polys[0], polys[1])
            a = rc.winding_raster(polys[0], 0, 0, 200 * rc.S, 200 * rc.S) != 0
            b = rc.winding_raster(polys[1], 0, 0, 200 * rc.S, 200 * rc.S) != 0
            est_i, est_u = (a & b).sum() / rc.S ** 2, (a | b).sum() / rc.S ** 2
            assert abs(inter - est_i) < 0.02 * max(est_i, 10.0), (polys[0].tolist(), polys[1].tolist(), inter, est_i)
            assert abs(iou - est_i / est_u) < 0.01, (polys[0].tolist(), polys[1].tolist(), iou, est_i / est_u)
    finally:
        rc.S = 4
