"""GPU: the polygon -> glyph crop glue against its oracle definition (crop spec v1,
oracle/postproc_oracle.c orc_crop_glyphs) — tiles bit-exact — and the one-call detect + read path end to end."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from ocr_rs_b200 import _ffi, pipeline, synth
    from ocr_rs_b200.char_recognition.model import Net
    from ocr_rs_b200.text_detection.model import resnet18
    from oracle import model_oracle as mo
    from oracle import postproc as pp
    return _ffi, pipeline, synth, Net, resnet18, mo, pp


def _random_box(rng, W, H):
    cx, cy = rng.uniform(0, W), rng.uniform(0, H)  # centres anywhere: boxes may leave the image (clamped sampling)
    w, h, a = rng.uniform(1, 260), rng.uniform(1, 90), rng.uniform(0, 2 * np.pi)
    if rng.random() < 0.3:
        w, h = h, w  # vertical text: the longer side is the left edge
    c, s = np.cos(a), np.sin(a)
    pts = np.array([[-w, -h], [w, -h], [w, h], [-w, h]]) / 2
    return np.round(pts @ np.array([[c, -s], [s, c]]).T + (cx, cy)).astype(np.int32)


def test_crop_tiles_bit_exact(env):
    _ffi, pipeline, synth, _, _, _, pp = env
    rng = np.random.default_rng(0)
    H, W = 416, 608
    img = rng.integers(0, 256, (H, W), dtype=np.uint8)
    img[100:300, 200:400] = synth.make_document_images(1, 200, 200, seed=1, n_boxes=4)[0]
    boxes = [_random_box(rng, W, H) for _ in range(300)]
    boxes += [np.array([(4, 10), (116, 10), (116, 38), (4, 38)]),       # 4 cells of exactly 28x28: identity
              np.array([(10, 10), (12, 10), (12, 11), (10, 11)]),       # fewer patch columns than cells
              np.array([(50, 50), (50, 50), (50, 50), (50, 50)]),       # degenerate box
              np.array([(-40, -30), (700, -30), (700, 500), (-40, 500)])]  # larger than the image
    for k in (1, 4, 7):
        got = pipeline.crop_glyphs(img, boxes, k).reshape(len(boxes), k, 784)
        for i, b in enumerate(boxes):
            want = pp.crop_glyphs(img, b, k)
            assert (got[i] == want).all(), (k, i, b.tolist())
    ident = pipeline.crop_glyphs(img, boxes[300:301], 4).reshape(4, 28, 28)
    assert (ident[0] == img[10:38, 4:32]).all() and (ident[3] == img[10:38, 88:116]).all()


def test_crop_very_wide_cell(env):
    """a cell wider than the shared-memory strip (1024 columns) takes the windowed path"""
    _ffi, pipeline, _, _, _, _, pp = env
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (64, 12000), dtype=np.uint8)
    boxes = [np.array([(5, 3), (11000, 9), (11000, 49), (5, 43)]), np.array([(100, 10), (3000, 10), (3000, 40), (100, 40)])]
    got = pipeline.crop_glyphs(img, boxes, 2).reshape(2, 2, 784)
    for i, b in enumerate(boxes):
        assert (got[i] == pp.crop_glyphs(img, b, 2)).all()


@pytest.mark.parametrize("device_images", [False, True])
def test_detect_and_read_end_to_end(env, device_images):
    """One call: detector -> post-processing -> crops of every kept polygon -> classes.  Checked stage by stage against the
    oracle run on the device's own map: same kept boxes -> same tiles (oracle crop) -> same classes (torch restatement),
    for host and device-resident images (several chunks with ramped host copies; the multi-group plans are covered by
    test_pipeline_group_and_chunk_knobs and the 1024-image test)."""
    torch = pytest.importorskip("torch")
    _ffi, pipeline, synth, Net, resnet18, mo, pp = env
    B, H, W, K = 70, 160, 160, 3
    wd = synth.make_detector_weights(0, "structured")
    wr = synth.make_rec_weights(1)
    imgs = synth.document_image_shard(0, B, H, W, unique=50)
    adj = np.ones((B, 2))
    adj[1::2] = (1.5, 0.75)
    det, rec = resnet18(wd, "bf16"), Net(wr)
    src = torch.from_numpy(imgs).cuda() if device_images else imgs
    res = pipeline.detect_and_read(det, rec, src, adj, K)
    plain, _ = pipeline.detect_and_recognize(det, rec, src, adj)
    for x, y in zip(res.arrays(), plain.arrays()):
        assert x.shape == y.shape and (x == y).all()
    n = len(res.all_scores)
    assert n > 30 and res.glyph_classes.shape == (n, K)
    prob = det.forward_t(imgs.reshape(B, 1, H, W))
    tiles = []
    for b in range(B):
        boxes = pp.kept_boxes_from_bitmap(prob[b, 0], pp.binarize(prob[b, 0], 0.6))
        assert len(boxes) == len(res.polygons[b])
        tiles += [pp.crop_glyphs(imgs[b], box, K) for box in boxes]
    tiles = np.concatenate(tiles)
    want = mo.rec_top1(mo.rec_forward(wr, tiles.astype(np.float32) / np.float32(255.0)))[0].reshape(n, K)
    assert (res.glyph_classes == want).all()
    assert len(np.unique(res.glyph_classes)) > 1  # the classes depend on the crops
