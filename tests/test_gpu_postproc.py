"""GPU parity: CUDA post-processing (through the C ABI) vs the oracle and the reference's own
golden vectors (metrics.rs:406-646).  Bar: bit-exact bitmaps / labels / chains / polygons,
scores equal to 1e-12 (both sides accumulate the same f32 values in f64)."""
import numpy as np
import pytest

import conftest as cf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from ocr_rs_b200 import polygon, synth
    from ocr_rs_b200.text_detection import metrics
    from oracle import postproc as pp
    return metrics, polygon, synth, pp


def test_binarize_kat_and_ties(api):
    metrics, _, synth, pp = api
    out = metrics.binarize(cf.KAT_BINARIZE_IN.astype(np.float32), 0.57)
    assert (out == cf.KAT_BINARIZE_OUT).all()
    t = np.float32(0.6)
    edge = np.array([t, np.nextafter(t, np.float32(1)), np.nextafter(t, np.float32(0)), 0.6000005, 0.5999995, np.nan, 1.0, 0.0], np.float32)
    assert metrics.binarize(edge, 0.6).tolist() == pp.binarize(edge, 0.6).tolist() == [0, 1, 0, 1, 0, 0, 1, 0]
    for n in (1, 3, 4, 5, 17, 1000003):  # ragged tails around the vector width
        p = np.random.default_rng(n).uniform(0.55, 0.65, size=n).astype(np.float32)
        assert (metrics.binarize(p, 0.6) == pp.binarize(p, 0.6)).all()


@pytest.mark.parametrize("pts,expected", cf.KAT_BOX_SCORES)
def test_box_score_kats(api, pts, expected):
    metrics = api[0]
    assert metrics.box_score_fast(cf.KAT_MAP_5x5, pts) == expected


def test_box_score_many_vertices(api):
    """No vertex limit (the reference has none): 300- and 1000-vertex polygons take the global-memory path."""
    metrics, _, synth, pp = api
    rng = np.random.default_rng(5)
    pred = rng.random((400, 500), dtype=np.float32)
    for n in (256, 257, 300, 1000):
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        r = rng.uniform(60, 190, n)
        pts = np.stack([250 + r * np.cos(ang), 200 + r * np.sin(ang)], 1).round().astype(np.int32)
        assert abs(metrics.box_score_fast(pred, pts) - pp.box_score(pred, pts)) <= 1e-12


def test_min_area_bounding_box_kat(api):
    metrics = api[0]
    box, sside = metrics.get_min_area_bounding_box(cf.KAT_MINRECT_IN)
    assert box.tolist() == [list(p) for p in cf.KAT_MINRECT_BOX]
    assert abs(sside - cf.KAT_MINRECT_SSIDE) < 1e-13


@pytest.mark.parametrize("adjust,expected", [((1.0, 1.0), cf.GOLDEN_POLYS_1X), ((2.0, 2.0), cf.GOLDEN_POLYS_2X)])
def test_get_polygons_from_bitmap_golden(api, gt55, adjust, expected):
    metrics = api[0]
    polys, scores = metrics.get_polygons_from_bitmap(gt55.astype(np.float32), gt55, adjust)
    assert [[tuple(int(v) for v in p) for p in poly] for poly in polys] == expected
    assert scores.tolist() == cf.GOLDEN_SCORES


def test_other_reference_bitmaps(api, gt_others):
    metrics, _, _, pp = api
    for name, bm in gt_others.items():
        pred = bm.astype(np.float32)
        got_p, got_s = metrics.get_polygons_from_bitmap(pred, bm, (1.0, 1.0))
        exp_p, exp_s = pp.polygons_from_bitmap(pred, bm, (1.0, 1.0))
        assert len(got_p) == len(exp_p) > 0, name
        for a, b in zip(got_p, exp_p):
            assert (a == b).all(), name
        assert (got_s == exp_s).all(), name


def test_ccl_labels_vs_scipy(api):
    metrics, _, synth, _ = api
    ndi = pytest.importorskip("scipy.ndimage")
    for seed, (h, w) in enumerate([(1, 1), (7, 5), (33, 65), (64, 64), (100, 130), (257, 513), (800, 800)]):
        bm = synth.make_random_bitmap(h, w, seed, density=0.25 + 0.05 * (seed % 5), smooth=seed % 3)
        lab, n = metrics.ccl_labels(bm)
        exp, n_exp = ndi.label(bm, structure=np.ones((3, 3)))
        assert n == n_exp and (lab == exp).all(), (h, w)
    # batched + degenerate
    bm = np.stack([synth.make_random_bitmap(96, 160, s, 0.4, 1) for s in range(3)] + [np.zeros((96, 160), np.uint8), np.ones((96, 160), np.uint8)])
    lab, n = metrics.ccl_labels(bm)
    for b in range(5):
        exp, n_exp = ndi.label(bm[b], structure=np.ones((3, 3)))
        assert n[b] == n_exp and (lab[b] == exp).all()


def test_find_contours_vs_oracle(api, gt55):
    metrics, _, synth, pp = api
    cases = [gt55] + [synth.make_random_bitmap(h, w, s, d, sm) for s, (h, w, d, sm) in enumerate(
        [(48, 64, 0.3, 0), (48, 64, 0.5, 1), (64, 48, 0.7, 2), (200, 300, 0.45, 2), (31, 17, 0.6, 0), (1, 9, 0.5, 0), (9, 1, 0.5, 0)])]
    cases += [np.zeros((16, 16), np.uint8), np.ones((16, 16), np.uint8)]
    total = 0
    for bm in cases:
        got, gt = metrics.find_contours(bm)
        exp, et = pp.find_contours(bm)
        assert len(got) == len(exp), bm.shape
        assert (gt == et).all()
        for a, b in zip(got, exp):
            assert a.shape == b.shape and (a == b).all()
        total += len(exp)
    assert total > 500


def test_approx_polygon_vs_oracle(api, gt55):
    metrics, _, synth, pp = api
    chains, _ = pp.find_contours(gt55)
    bm = synth.make_random_bitmap(200, 300, 11, 0.45, 2)
    chains += [c for c in pp.find_contours(bm)[0] if len(c) >= 1][:200]
    for c in chains:
        assert metrics.approx_polygon(c).tolist() == pp.dp_polygon(c).tolist()


def test_expand_polygon_vs_oracle(api, gt55):
    metrics, polygon, synth, pp = api
    polys = [pp.dp_polygon(c) for c in pp.find_contours(gt55)[0]]
    bm = synth.make_random_bitmap(300, 400, 5, 0.5, 3)
    polys += [pp.dp_polygon(c) for c in pp.find_contours(bm)[0]]
    polys = [p for p in polys if len(p) >= 4][:150]
    polys.append(np.array([(10, 10), (10, 50), (50, 50), (50, 10)]))
    n_some = n_exact = 0
    for p in polys:
        exp = pp.expand_polygon(p, 2.0)
        got = polygon.expand_polygon(p, 2.0)
        if exp is None:
            assert got is None
        else:
            assert got is not None and got.tolist() == exp.tolist()
            n_some += 1
            be, se = pp.min_area_bounding_box(exp)
            bg, sg = metrics.get_min_area_bounding_box(exp)
            # f64 trig: the device evaluates atan2/sin/cos correctly rounded (csrc/dd_math.cuh), glibc
            # (the oracle, like the reference's libm) misrounds ~0.1 % of calls by one ulp, which can
            # flip the outward floor/ceil when a rotated coordinate is an exact integer
            # On this fixed set every box is bit-identical; a failure here lists the polygon so that a glibc
            # misround can be told from a real difference (tests/test_dd_math.py arbitrates with a 60-digit series).
            assert bg.tolist() == be.tolist() and abs(sg - se) <= 1e-12 * max(1.0, se), (exp.tolist(), bg.tolist(), be.tolist(), sg, se)
            n_exact += 1
    assert n_some > 20 and n_exact == n_some, (n_some, n_exact)
    print(f"min-area-rect hook: {n_exact}/{n_some} boxes bit-identical to the oracle")
    assert polygon.expand_polygon([(0, 0), (10, 0), (20, 0), (10, 0)], 2.0) is None


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4, 5, 6])
def test_expand_polygon_random_shapes(api, kind):
    """polygon.rs:13-56 on 7 x 1500 random polygons (rectangles, concave / spiky stars, near-collinear slivers,
    notched shapes whose notch the expansion closes, mild and chaotic self-intersections): the CUDA walk must
    return the oracle's vertices, and — independently of both — the region it encloses must be the positive-
    winding region of the raw offset path (oracle/region_check.c: a scanline winding rasteriser that shares
    nothing with either implementation)."""
    metrics, polygon, synth, pp = api
    from oracle import region_check as rc
    rng = np.random.default_rng(1000 + kind)
    n_checked = 0
    for i in range(1500):
        poly = rc.random_dp_polygon(rng, kind)
        exp, d = pp.clip_polygon(poly, 2.0, False, True)
        got = polygon.expand_polygon(poly, 2.0)
        if exp is None:
            assert got is None, (i, poly.tolist())
            continue
        assert got is not None and got.tolist() == exp.tolist(), (i, poly.tolist())
        if kind < 6 and i % 5 == 0:
            r = rc.check_multires(pp.offset_raw(poly, d), got)
            assert r["ok"], (i, poly.tolist(), r)
            n_checked += 1
    assert kind == 6 or n_checked >= 250


def _compare_maps(metrics, pp, prob, adjust):
    """prob [B,H,W]; full get_boxes_and_box_scores vs the oracle, image by image."""
    B = prob.shape[0]
    res = metrics.get_boxes_and_box_scores(prob.reshape(B, 1, *prob.shape[1:]), adjust)
    n = 0
    for b in range(B):
        exp_p, exp_s, st = pp.polygons_from_bitmap(prob[b], pp.binarize(prob[b], 0.6), tuple(adjust[b]), return_stats=True)
        assert res.stats[b].tolist() == st.tolist(), (b, res.stats[b], st)
        assert len(res.polygons[b]) == len(exp_p)
        for a, e in zip(res.polygons[b], exp_p):
            assert a.shape == e.shape and (a == e).all()
        assert np.allclose(res.scores[b], exp_s, rtol=0, atol=1e-12, equal_nan=True)  # NaN = empty mask (0/0) on both sides
        n += len(exp_p)
    return n


def test_full_postproc_blob_maps(api):
    metrics, _, synth, pp = api
    prob = np.stack([synth.make_blob_prob_map(800, 800, 60, seed=s) for s in (4, 5, 6)])
    adjust = np.array([[1.0, 1.0], [800 / 300, 533 / 200], [0.5, 2.0]])
    assert _compare_maps(metrics, pp, prob, adjust) > 100


def test_full_postproc_noise_and_edges(api):
    metrics, _, synth, pp = api
    rng = np.random.default_rng(9)
    # salt-and-pepper map (thousands of tiny contours, holes) and frame-touching blobs
    noise = rng.uniform(0.25, 0.85, size=(1, 320, 320)).astype(np.float32)
    assert _compare_maps(metrics, pp, noise, np.ones((1, 2))) >= 0
    touching = synth.make_random_bitmap(256, 384, 3, 0.55, 3).astype(np.float32)[None] * np.float32(0.9)
    assert _compare_maps(metrics, pp, touching, np.ones((1, 2))) > 0
    empty = np.zeros((2, 64, 96), np.float32)
    assert _compare_maps(metrics, pp, empty, np.ones((2, 2))) == 0
    full = np.ones((1, 64, 96), np.float32)
    _compare_maps(metrics, pp, full, np.ones((1, 2)))


def test_postproc_stress_4096(api):
    """BASELINE config 5: 4096x4096 map, ~10k components."""
    metrics, _, synth, pp = api
    prob = synth.make_blob_prob_map(4096, 4096, 9000, seed=4, near_thresh=4096, max_w=48, max_h=24)[None]
    n = _compare_maps(metrics, pp, prob, np.ones((1, 2)))
    assert n > 7000


def test_image_ops_vs_oracle(api, preprocessed):
    _, _, _, pp = api
    from ocr_rs_b200 import image_ops
    for name in ("img55", "img545"):
        src = preprocessed["src_" + name]
        got, ax, ay = image_ops.preprocess_image(src, (800, 800))
        exp, ex, ey = pp.preprocess(src, 800, 800)
        assert (ax, ay) == (ex, ey)
        assert (got == exp).all(), name
    rng = np.random.default_rng(0)
    for (h, w), (W, H) in [((37, 53), (64, 64)), ((1000, 300), (160, 96)), ((64, 64), (64, 64)), ((5, 400), (128, 32))]:
        src = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        got, ax, ay = image_ops.preprocess_image(src, (W, H))
        exp, ex, ey = pp.preprocess(src, W, H)
        assert (ax, ay) == (ex, ey) and (got == exp).all(), ((h, w), (W, H))
    img = rng.integers(0, 256, size=(33, 47), dtype=np.uint8)
    assert (image_ops.convert_image_to_tensor(img) == img.astype(np.float32)).all()
    t = rng.uniform(0, 1, size=(33, 47)).astype(np.float32)
    assert (image_ops.convert_tensor_to_image(t, 255.0) == (t * np.float32(255.0)).astype(np.uint8)).all()
    assert (image_ops.load_image_as_tensor(img) == (img.astype(np.float32) / np.float32(255.0)).reshape(1, -1)).all()


def test_preprocess_batch_vs_oracle(api, preprocessed):
    """ocrb_preprocess_rgba_batch (one fused launch, the u8 intermediate in shared memory only) is bit-identical to the
    oracle's two-pass resize + luma + pad per image: the reference's JPEG sources, identity, up- and down-scaling,
    portrait / landscape, and a down-scaling factor beyond the fused kernel's tile span (fallback path)."""
    from ocr_rs_b200 import image_ops
    _, _, synth, pp = api
    rng = np.random.default_rng(3)
    imgs = [preprocessed["src_img55"], preprocessed["src_img545"]]
    for (h, w) in ((800, 800), (1600, 1600), (533, 800), (2400, 1100), (97, 1311), (64, 64), (3, 5)):
        imgs.append(rng.integers(0, 256, (h, w, 4), dtype=np.uint8))
    got, adj = image_ops.preprocess_images(imgs, (800, 800))
    for i, im in enumerate(imgs):
        want, ax, ay = pp.preprocess(im, 800, 800)
        assert (got[i] == want).all(), (i, im.shape)
        assert (adj[i] == (ax, ay)).all()
        one, ax1, ay1 = image_ops.preprocess_image(im, (800, 800))
        assert (one == want).all()
    # beyond the tile span: 12 x down-scaling of a wide strip
    wide = [rng.integers(0, 256, (80, 9600, 4), dtype=np.uint8), imgs[2]]
    got, _ = image_ops.preprocess_images(wide, (800, 800))
    for i, im in enumerate(wide):
        assert (got[i] == pp.preprocess(im, 800, 800)[0]).all()
