"""CPU: the library's own offset / union source — ocr_rs_b200/csrc/geometry.cu clipper_offset_raw + union_positive, the
functions the device unclip kernel runs, compiled __host__ __device__ and reached through the host entry point
ocrb_clip_polygon (polygon.rs:13-49) — against
  * the oracle (oracle/postproc_oracle.c orc_clip_polygon) bit for bit on thousands of random Douglas-Peucker-like polygons,
    both offset signs;
  * the reference's own fixtures: gt_shrinked_img{55,224,494}.png regenerated from its ground-truth polygon files through
    the LIBRARY's shrink_polygon (generate_gt_and_mask_images, image_ops.rs:222-277; pinned by image_ops.rs:805-1008);
  * the reference's golden expanded polygons (metrics.rs:510-646 go through the same code on the device).
No device needed: this is the a9 source held to its pins in the CPU suite."""
import os

import numpy as np
import pytest

import conftest as cf
from ocr_rs_b200 import polygon
from oracle import postproc as pp
from oracle import region_check as rc


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("shrink", [False, True])
def test_library_source_equals_oracle(kind, shrink):
    rng = np.random.default_rng(7000 + 10 * kind + int(shrink))
    n_some = 0
    for _ in range(600):
        poly = rc.random_dp_polygon(rng, kind)
        want, dw = pp.clip_polygon(poly, 0.75 if shrink else 2.0, shrink, True)
        got, dg = polygon.clip_polygon(poly, 0.75 if shrink else 2.0, shrink, True)
        assert dg == dw
        assert (got is None) == (want is None), (poly.tolist(), got, want)
        if want is not None:
            assert got.shape == want.shape and (got == want).all(), (poly.tolist(), got.tolist(), want.tolist())
            n_some += 1
    assert n_some > 100


def _gt_case(name):
    z = np.load(os.path.join(cf.GOLDEN, "text_det_gts.npz"))
    counts, pts = z[name + "_counts"], z[name + "_points"]
    polys, o = [], 0
    for c in counts:
        polys.append(pts[o:o + c])
        o += c
    ax, ay = z[name + "_resized"] / z[name + "_orig"]
    return polys, float(ax), float(ay)


@pytest.mark.parametrize("name", ["img55", "img224", "img494"])
def test_reference_shrinked_maps_from_the_library_source(name, gt55, gt_others):
    polys, ax, ay = _gt_case(name)
    canvas = np.zeros((800, 800), np.uint8)
    for poly in polys:
        # image_ops.rs:253-262: scale by the adjust factors, truncate to integers
        vals = np.stack([(poly[:, 0].astype(np.float64) * ax).astype(np.int32), (poly[:, 1].astype(np.float64) * ay).astype(np.int32)], 1)
        sh = polygon.shrink_polygon(vals, 0.75)  # 1 - 0.5^2, image_ops.rs:265
        assert sh is not None
        pp.draw_polygon(canvas, sh, 255)
    want = gt55 if name == "img55" else gt_others[name]
    assert ((canvas > 0) == (want > 0)).all()


def test_degenerate_and_capacity():
    assert polygon.clip_polygon([(5, 5), (5, 5), (5, 5), (5, 5)], 2.0, False) is None
    assert polygon.clip_polygon([(0, 0), (10, 0), (20, 0), (10, 0)], 2.0, False) is None  # zero area: the reference's panic (D11)
    sq = polygon.clip_polygon([(10, 10), (20, 10), (20, 20), (10, 20)], 2.0, False)
    assert sorted(map(tuple, sq.tolist())) == [(5, 5), (5, 25), (25, 5), (25, 25)]
    import ctypes as C
    from ocr_rs_b200 import _ffi
    p = np.array([(10, 10), (20, 10), (20, 20), (10, 20)], np.int32)
    out = np.zeros((2, 2), np.int32)
    n = C.c_int(0)
    assert _ffi.lib().ocrb_clip_polygon(_ffi.ptr(p), 4, 2.0, 0, _ffi.ptr(out), 2, C.byref(n), None) == -3  # OCRB_ERR_CAPACITY


def _minrect_host(points):
    import ctypes as C
    from ocr_rs_b200 import _ffi
    p = np.ascontiguousarray(np.asarray(points, np.int32).reshape(-1, 2))
    box = np.empty((4, 2), np.int32)
    s = C.c_double(0.0)
    _ffi.check(_ffi.lib().ocrb_debug_min_area_bounding_box_host(_ffi.ptr(p), len(p), _ffi.ptr(box), C.byref(s)))
    return box, s.value


def test_min_area_rect_source_on_the_host(gt55):
    """get_min_area_bounding_box (metrics.rs:133-148): the function the unclip kernel runs (geometry.cu
    min_area_bounding_box, compiled __host__ __device__) on the host — the reference's known answer (metrics.rs:406-424)
    exactly, and the oracle on the fixed polygon set of the GPU test bit for bit.  On random polygons the two may differ
    where glibc's atan2 / sin / cos (oracle, like the reference's libm) misround by an ulp and an outward floor / ceil
    flips (the library evaluates them correctly rounded, csrc/dd_math.cuh; tests/test_dd_math.py arbitrates): a statistic."""
    from ocr_rs_b200 import synth
    box, sside = _minrect_host(cf.KAT_MINRECT_IN)
    assert box.tolist() == [list(p) for p in cf.KAT_MINRECT_BOX]
    assert abs(sside - cf.KAT_MINRECT_SSIDE) < np.finfo(np.float64).eps
    polys = [pp.dp_polygon(c) for c in pp.find_contours(gt55)[0]]
    bm = synth.make_random_bitmap(300, 400, 5, 0.5, 3)
    polys += [pp.dp_polygon(c) for c in pp.find_contours(bm)[0]]
    polys = [p for p in polys if len(p) >= 4][:150]
    n = 0
    for p in polys:
        exp = polygon.clip_polygon(p, 2.0, False)
        if exp is None:
            continue
        be, se = pp.min_area_bounding_box(exp)
        bg, sg = _minrect_host(exp)
        assert bg.tolist() == be.tolist() and abs(sg - se) <= 1e-12 * max(1.0, se), (exp.tolist(), bg.tolist(), be.tolist())
        n += 1
    assert n > 20
    rng = np.random.default_rng(99)
    same = total = 0
    for kind in range(6):
        for _ in range(300):
            exp = polygon.clip_polygon(rc.random_dp_polygon(rng, kind), 2.0, False)
            if exp is None:
                continue
            be, se = pp.min_area_bounding_box(exp)
            bg, sg = _minrect_host(exp)
            total += 1
            same += bg.tolist() == be.tolist() and abs(sg - se) <= 1e-12 * max(1.0, se)
            assert np.abs(bg - be).max() <= 1  # a flipped floor / ceil moves a corner by one pixel at most
    assert total > 1000 and same >= 0.99 * total, (same, total)


def test_douglas_peucker_kernel_statements_on_the_host(gt55):
    """approximate_polygon_dp as metrics.rs:87-95 uses it: the device kernel's own statements (contours.cu approx_dp_kernel,
    the block between the [approx-dp-body] markers, repeated verbatim in the host hook — checked textually here) against the
    oracle on every border of the reference's gt_shrinked_img55 map and of random bitmaps, exactly."""
    import ctypes as C
    import re
    from ocr_rs_b200 import _ffi, synth
    src = open(os.path.join(os.path.dirname(cf.GOLDEN), "..", "ocr_rs_b200", "csrc", "contours.cu")).read()
    blocks = re.findall(r"// \[approx-dp-body-begin\][^\n]*\n(.*?)// \[approx-dp-body-end\]", src, re.S)
    assert len(blocks) == 2
    norm = [[l.strip() for l in b.splitlines() if l.strip()] for b in blocks]
    assert norm[0] == norm[1] and len(norm[0]) > 30, "the host hook's body is no longer the kernel's"

    def dp_host(chain):
        p = np.ascontiguousarray(np.asarray(chain, np.int32).reshape(-1, 2))
        out = np.empty((len(p) + 1, 2), np.int32)
        n = C.c_int64(0)
        _ffi.check(_ffi.lib().ocrb_debug_approx_polygon_host(_ffi.ptr(p), len(p), _ffi.ptr(out), len(out), C.byref(n)))
        return out[: n.value]

    maps = [gt55] + [synth.make_random_bitmap(200, 260, 5 + s, 0.5, s) for s in range(4)]
    n_chains = 0
    for bm in maps:
        for chain in pp.find_contours(bm)[0]:
            want = pp.dp_polygon(chain)
            got = dp_host(chain)
            assert got.shape == want.shape and (got == want).all(), (chain.tolist(), got.tolist(), want.tolist())
            n_chains += 1
    assert n_chains > 300
    # single pixels and two-pixel borders
    assert dp_host([(7, 9)]).tolist() == pp.dp_polygon(np.array([(7, 9)])).tolist()
