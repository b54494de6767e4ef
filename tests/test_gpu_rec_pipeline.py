"""GPU parity: glyph recognition (char_recognition/model.rs:12-39, mod.rs:53-56) and the
batched detect+recognize pipeline (text_detection/mod.rs:46-67, :188-204) through the C ABI."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from ocr_rs_b200 import _ffi, synth
    from ocr_rs_b200.char_recognition.model import Net
    from ocr_rs_b200.text_detection.model import resnet18
    from oracle import model_oracle as mo
    from oracle import postproc as pp
    return _ffi, synth, Net, resnet18, mo, pp


@pytest.mark.parametrize("kind,n", [("noise", 4096), ("strokes", 1000), ("noise", 1), ("strokes", 67)])
def test_rec_logits_and_argmax(env, kind, n):
    """BASELINE config 2 (28x28 glyphs): logits <= 1e-4, class argmax identical."""
    _ffi, synth, Net, _, mo, _ = env
    w = synth.make_rec_weights(1)
    g = synth.make_glyphs(n, 1, kind)
    net = Net(w)
    logits, argmax, prob = net.predict(g)  # u8 path (x/255 fused)
    x = g.astype(np.float32) / np.float32(255.0)
    ref = mo.rec_forward(w, x)
    want, want_p = mo.rec_top1(ref)
    ref = ref.numpy()
    assert np.abs(logits - ref).max() <= 1e-4
    top2 = np.sort(ref, -1)[:, -2:]
    ties = int(((top2[:, 1] - top2[:, 0]) < 1e-4).sum())
    mism = int((argmax != want).sum())
    print(f"rec {kind} n={n}: max|dlogit| = {np.abs(logits - ref).max():.2e}, top-2 gap < 1e-4 on {ties} glyphs, {mism} argmax mismatches")
    assert mism == 0
    assert np.abs(prob - want_p).max() <= 1e-6
    # the f32 entry point (conv1 on CUDA cores: arbitrary floats cannot use the exact-u8 tensor-core form) agrees with the
    # u8 one to fp32 summation-order noise and gives the same classes
    l2, a2, _ = net.predict(x)
    assert np.abs(l2 - logits).max() <= 2e-6 and (a2 == argmax).all()
    assert np.abs(l2 - ref).max() <= 1e-4


@pytest.mark.parametrize("knobs", [{"OCRB_REC": "fp32"}, {"OCRB_REC_CONV1": "cuda"}, {"OCRB_REC_CONV2": "smem"},
                                   {"OCRB_REC_CONV1": "cuda", "OCRB_REC_CONV2": "smem"}])
def test_rec_alternate_paths(knobs):
    """Knobs of the glyph net (read once per process): OCRB_REC=fp32 the all-CUDA-core fp32 path, OCRB_REC_CONV1=cuda conv1
    on CUDA cores, OCRB_REC_CONV2=smem conv2 with the shared-memory builder warps instead of TMA boxes.  Same classes,
    logits within 1e-5 of the oracle."""
    import os
    import subprocess
    import sys
    script = r"""
import numpy as np
from ocr_rs_b200 import synth
from ocr_rs_b200.char_recognition.model import Net
from oracle import model_oracle as mo
w = synth.make_rec_weights(1)
g = synth.make_glyphs(777, 1, "strokes")
logits, argmax, _ = Net(w).predict(g)
ref = mo.rec_forward(w, g.astype(np.float32) / np.float32(255.0)).numpy()
print("WORST", float(np.abs(logits - ref).max()), int((argmax != ref.argmax(-1)).sum()))
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""), **knobs)
    out = subprocess.run([sys.executable, "-c", script], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    worst, mism = [l for l in out.stdout.splitlines() if l.startswith("WORST")][-1].split()[1:]
    assert float(worst) <= 1e-5 and int(mism) == 0


def test_rec_varstore_aliases_and_errors(env):
    _ffi, synth, Net, _, _, _ = env
    from ocr_rs_b200 import OcrbError, utils
    w = synth.make_rec_weights(3)
    aliased = {a: w[n] for (n, _), a in zip(synth.REC_CANONICAL, synth.REC_VARSTORE_ALIASES)}
    g = synth.make_glyphs(16, 2, "strokes")
    assert (Net(w).predict(g)[1] == Net(aliased).predict(g)[1]).all()
    # the other creation order (linear: weight before bias), and under a path prefix: the suffix is not interpreted
    other = {a: w[n] for (n, _), a in zip(synth.REC_CANONICAL, synth.REC_VARSTORE_ALIASES_WEIGHT_FIRST_LINEAR)}
    assert (Net(w).predict(g)[0] == Net(other).predict(g)[0]).all()
    prefixed = {"net." + a: v for a, v in other.items()}
    assert (Net(w).predict(g)[0] == Net(prefixed).predict(g)[0]).all()
    wrong = dict(aliased)
    wrong["weight__9"] = np.zeros(7, np.float32)
    with pytest.raises(OcrbError):
        Net(wrong)
    bad = dict(w)
    del bad["fc2.bias"]
    with pytest.raises(OcrbError):
        Net(bad)
    assert "".join(utils.class_to_char(i) for i in range(62)) == utils.VALUES


def _run_pipeline(_ffi, det, rec, imgs, adj, glyphs):
    B, H, W = imgs.shape
    am = np.empty(len(glyphs), np.int32)
    h = _ffi.c_p()
    _ffi.check(_ffi.lib().ocrb_detect_and_recognize(det._h, rec._h, _ffi.ptr(imgs), _ffi.ptr(adj), B, H, W, None,
                                                    _ffi.ptr(glyphs), len(glyphs), _ffi.ptr(am), C.byref(h)))
    return _ffi.Polygons(h), am


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_pipeline_structured_weights(env, mode):
    """BASELINE config 3 shape at reduced batch: document images, structured head:
    post-processing of the device's own map is identical to the oracle's; against the oracle's
    own end-to-end path (torch map -> C post-proc) polygons agree at IoU >= 0.99 wherever no
    pixel of the map sits within tolerance of a decision threshold."""
    _ffi, synth, Net, resnet18, mo, pp = env
    B, H, W = 3, 800, 800
    wd = synth.make_detector_weights(0, "structured")
    wr = synth.make_rec_weights(1)
    imgs = synth.make_document_images(B, H, W, seed=3)
    glyphs = synth.make_glyphs(256, 4, "strokes")
    adj = np.array([[1.0, 1.0], [800 / 300, 533 / 200], [2.0, 0.5]])
    det, rec = resnet18(wd, mode), Net(wr)
    res, am = _run_pipeline(_ffi, det, rec, imgs, adj, glyphs)
    prob = det.forward_t(imgs.reshape(B, 1, H, W))
    total = 0
    for b in range(B):
        exp_p, exp_s = pp.polygons_from_bitmap(prob[b, 0], pp.binarize(prob[b, 0], 0.6), tuple(adj[b]))
        assert len(res.polygons[b]) == len(exp_p)
        for a, e in zip(res.polygons[b], exp_p):
            assert a.shape == e.shape and (a == e).all()
        assert np.allclose(res.scores[b], exp_s, rtol=0, atol=1e-12, equal_nan=True)  # NaN = empty mask (0/0) on both sides
        total += len(exp_p)
    assert total > 0
    want = mo.rec_top1(mo.rec_forward(wr, glyphs.astype(np.float32) / np.float32(255.0)))[0]
    assert (am == want).all()
    # end to end vs the oracle's own map: every decidable polygon (tests/e2e_compare.py states the rule) must
    # have a partner at IoU >= 0.99, in both directions, in BOTH arithmetic modes
    import e2e_compare
    ref = mo.detector_forward(wd, imgs.reshape(B, 1, H, W).astype(np.float32)).numpy()
    matched = undecidable = 0
    for b in range(B):
        r = e2e_compare.compare(ref[b, 0], res.polygons[b], res.scores[b], adj[b], 1e-2 if mode == "bf16" else 1e-4, 64.0)
        assert not r["failures"], (mode, b, r["failures"][:3])
        matched += r["matched"]
        undecidable += r["undecidable"]
    print(f"pipeline {mode}: {total} polygons from the device map; vs oracle end-to-end {matched} matched at IoU>=0.99, {undecidable} undecidable")
    assert matched >= 40 and undecidable <= matched // 4, (matched, undecidable)


def test_pipeline_chunking_and_device_pointers(env):
    """More images than one chunk; device-resident inputs give the same result as host inputs."""
    torch = pytest.importorskip("torch")
    _ffi, synth, Net, resnet18, _, _ = env
    B, H, W = 37, 96, 128
    wd = synth.make_detector_weights(0, "structured")
    imgs = synth.make_document_images(B, H, W, seed=11, n_boxes=3)
    glyphs = synth.make_glyphs(8, 4, "strokes")
    adj = np.ones((B, 2))
    det, rec = resnet18(wd, "bf16"), Net(synth.make_rec_weights(1))
    res, am = _run_pipeline(_ffi, det, rec, imgs, adj, glyphs)
    res1 = [_run_pipeline(_ffi, det, rec, imgs[b:b + 1], adj[b:b + 1], glyphs)[0] for b in range(B)]
    for b in range(B):
        assert len(res.polygons[b]) == len(res1[b].polygons[0])
        for a, e in zip(res.polygons[b], res1[b].polygons[0]):
            assert (a == e).all()
    dimgs = torch.from_numpy(imgs).cuda()
    h = _ffi.c_p()
    _ffi.check(_ffi.lib().ocrb_detect_and_recognize(det._h, None, _ffi.ptr(dimgs), _ffi.ptr(adj), B, H, W, None, None, 0, None, C.byref(h)))
    res2 = _ffi.Polygons(h)
    assert (res2.xy == res.xy).all() and (res2.image_offsets == res.image_offsets).all()


def test_pipeline_sharding_invariance(env):
    """BASELINE config 4 property at reduced size: the polygons of an image do not depend on
    how the index range is cut (one call over 200 images == two shards == single-image calls),
    (round 1 cut such a batch into two groups; see test_pipeline_group_and_chunk_knobs for the multi-group plans)."""
    _ffi, synth, Net, resnet18, _, _ = env
    B, H, W = 200, 160, 160
    wd = synth.make_detector_weights(0, "structured")
    imgs = synth.document_image_shard(0, B, H, W, unique=50)
    # smaller boxes for the small frame
    glyphs = synth.make_glyphs(4, 4, "strokes")
    adj = np.ones((B, 2))
    det, rec = resnet18(wd, "bf16"), Net(synth.make_rec_weights(1))
    whole, _ = _run_pipeline(_ffi, det, rec, imgs, adj, glyphs)
    a, _ = _run_pipeline(_ffi, det, rec, imgs[:77], adj[:77], glyphs)
    b, _ = _run_pipeline(_ffi, det, rec, imgs[77:], adj[77:], glyphs)
    both = _ffi.Polygons.concat([a, b])
    for x, y in zip(whole.arrays(), both.arrays()):
        assert x.shape == y.shape and (x == y).all()
    for i in (0, 63, 64, 127, 128, 199):
        one, _ = _run_pipeline(_ffi, det, rec, imgs[i:i + 1], adj[i:i + 1], glyphs)
        assert len(one.polygons[0]) == len(whole.polygons[i])
        for p, q in zip(one.polygons[0], whole.polygons[i]):
            assert (p == q).all()
    assert sum(len(p) for p in whole.polygons) > 0


def test_config4_full_size_shard_invariance(env):
    """BASELINE config 4 at full size (1024 images 800x800, the bench workload): one call over the whole
    range == the concatenation of 8 contiguous shards of 128 (what 8 ranks compute) == single-image calls
    at sampled indices; recognised glyph classes do not depend on the call either."""
    _ffi, synth, Net, resnet18, _, _ = env
    B, H, W = 1024, 800, 800
    wd = synth.make_detector_weights(0, "structured")
    imgs = synth.document_image_shard(0, B, H, W)
    glyphs = synth.make_glyphs(64, 1, "strokes")
    adj = np.ones((B, 2))
    det, rec = resnet18(wd, "bf16"), Net(synth.make_rec_weights(1))
    whole, am = _run_pipeline(_ffi, det, rec, imgs, adj, glyphs)
    parts = []
    for r in range(8):
        p, am_r = _run_pipeline(_ffi, det, rec, imgs[r * 128:(r + 1) * 128], adj[r * 128:(r + 1) * 128], glyphs)
        parts.append(p)
        assert (am_r == am).all()
    both = _ffi.Polygons.concat(parts)
    for x, y in zip(whole.arrays(), both.arrays()):
        assert x.shape == y.shape and (x == y).all()
    for i in (0, 127, 128, 500, 1023):
        one, _ = _run_pipeline(_ffi, det, rec, imgs[i:i + 1], adj[i:i + 1], glyphs)
        assert len(one.polygons[0]) == len(whole.polygons[i])
        for p, q in zip(one.polygons[0], whole.polygons[i]):
            assert (p == q).all()
    n = sum(len(p) for p in whole.polygons)
    print(f"config 4 full size: {n} polygons over {B} images")
    assert n > 10000
    # the bench workload against the oracle's own end-to-end path (torch-CPU map -> C post-processing) on 32
    # images spread over all chunks and post-processing groups of the 1024-image call
    import e2e_compare
    _, _, _, _, mo, _ = env
    matched = undecidable = 0
    for i in np.linspace(0, B - 1, 32).round().astype(int):
        ref = mo.detector_forward(wd, imgs[i].reshape(1, 1, H, W).astype(np.float32)).numpy()[0, 0]
        r = e2e_compare.compare(ref, whole.polygons[i], whole.scores[i], adj[i], 1e-2, 64.0)
        assert not r["failures"], (int(i), r["failures"][:3])
        matched += r["matched"]
        undecidable += r["undecidable"]
    print(f"config 4 bench images vs oracle end-to-end: {matched} polygons matched at IoU>=0.99, {undecidable} undecidable, 0 failures")
    assert matched >= 500 and undecidable <= matched // 4, (matched, undecidable)


def test_sharded_entry_point(env):
    """ocrb_detect_and_recognize_sharded (one host thread per device inside the library): the whole
    batch's polygons and glyph classes are those of the single-device call, on one device and — when the box has
    them — on two and on all devices; more devices than images is legal."""
    torch = pytest.importorskip("torch")
    _ffi, synth, Net, resnet18, _, _ = env
    from ocr_rs_b200 import OcrbError, sharding
    B, H, W = 45, 160, 160
    wd = synth.make_detector_weights(0, "structured")
    wr = synth.make_rec_weights(1)
    imgs = synth.document_image_shard(0, B, H, W, unique=50)
    glyphs = synth.make_glyphs(4 * B + 3, 4, "strokes")
    adj = np.ones((B, 2))
    adj[::3] = (2.0, 0.5)
    det, rec = resnet18(wd, "bf16"), Net(wr)
    want, want_am = _run_pipeline(_ffi, det, rec, imgs, adj, glyphs)
    n_dev = torch.cuda.device_count()
    for devices in ([0], list(range(min(2, n_dev))), list(range(n_dev))):
        sh = sharding.Shards(devices, wd, wr, "bf16")
        got, am = sh.detect_and_recognize(imgs, adj, glyphs)
        for x, y in zip(want.arrays(), got.arrays()):
            assert x.shape == y.shape and (x == y).all(), devices
        assert (am == want_am).all()
        # a batch smaller than the device list
        got1, am1 = sh.detect_and_recognize(imgs[:1], adj[:1], glyphs[:2])
        assert (got1.xy == _run_pipeline(_ffi, det, rec, imgs[:1], adj[:1], glyphs[:2])[0].xy).all() and (am1 == want_am[:2]).all()
        assert sh.launch_count > 0
        sh.close()
    with pytest.raises(OcrbError):
        sharding.Shards([0, 0], wd, wr)
    with pytest.raises(OcrbError):
        sh = sharding.Shards([0], wd, None)
        sh.detect_and_recognize(torch.from_numpy(imgs).cuda(), adj)  # device pointers cannot feed several devices


@pytest.mark.gpu
def test_pipeline_group_and_chunk_knobs():
    """Batching must not change results: the default plan (one post-processing group for a batch of up to 256 images),
    small groups / chunks (OCRB_GROUP=32, OCRB_CHUNK=16: three groups, ramped host copies) and round 1's two-group rule
    (OCRB_GROUP_SPLIT=1) give the same polygons, scores and glyph classes.  The knobs are read once per process, so every
    plan runs in its own interpreter and prints a digest of the result arrays."""
    import os
    import subprocess
    import sys
    script = r"""
import hashlib
import numpy as np
from ocr_rs_b200 import pipeline, synth
from ocr_rs_b200.char_recognition.model import Net
from ocr_rs_b200.text_detection.model import resnet18
B, H, W, K = 70, 160, 160, 3
imgs = synth.document_image_shard(0, B, H, W, unique=50)
adj = np.ones((B, 2))
adj[1::2] = (1.5, 0.75)
det, rec = resnet18(synth.make_detector_weights(0, "structured"), "bf16"), Net(synth.make_rec_weights(1))
res = pipeline.detect_and_read(det, rec, imgs, adj, K)
h = hashlib.sha1()
for a in list(res.arrays()) + [res.glyph_classes]:
    h.update(np.ascontiguousarray(a).tobytes())
print("DIGEST", h.hexdigest(), len(res.all_scores))
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = []
    for knobs in ({}, {"OCRB_GROUP": "32", "OCRB_CHUNK": "16"}, {"OCRB_GROUP_SPLIT": "1"}):
        env = dict(os.environ, **knobs)
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        out = subprocess.run([sys.executable, "-c", script], env=env, cwd=root, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        line = [l for l in out.stdout.splitlines() if l.startswith("DIGEST")][-1].split()
        assert int(line[2]) > 30
        digests.append(line[1])
        print(knobs, line[1], line[2])
    assert digests[0] == digests[1] == digests[2]


def test_run_prediction_topk(env):
    """run_prediction (char_recognition/mod.rs:39-68) through the module mirror: one 28x28 luma glyph -> load_image_as_tensor ->
    Net -> softmax in f64 -> utils::topk.  k = 1 is the reference's call; k = 3 exercises the utils::topk mirror.  Checked
    against the torch restatement of the same steps."""
    import torch
    from ocr_rs_b200 import char_recognition, utils
    _ffi, synth, Net, resnet18, mo, _ = env
    wr = synth.make_rec_weights(1)
    for g in synth.make_glyphs(3, 5, "strokes").reshape(-1, 28, 28):
        logits = mo.rec_forward(wr, g.reshape(1, 784).astype(np.float32) / np.float32(255.0))
        p = torch.softmax(logits.to(torch.float64), -1)[0].numpy()
        want = utils.topk(p, 3)
        ch, prob = char_recognition.run_prediction(g, wr)
        assert ch == want[0][0] and abs(prob - want[0][1]) <= 1e-6
        got = char_recognition.run_prediction(g, wr, k=3)
        assert [c for c, _ in got] == [c for c, _ in want]
        assert all(abs(a - b) <= 1e-6 for (_, a), (_, b) in zip(got, want))
