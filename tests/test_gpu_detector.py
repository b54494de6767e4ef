"""GPU parity: detector forward (model.rs:65-156) through the C ABI vs the torch-CPU oracle.
Tolerances are the north_star's: probability maps within 1e-4 abs in FP32 mode, 1e-2 in BF16
mode.  The oracle forward is "parity unpinned" against ocr-rs itself (oracle/model_oracle.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 1e-2}


@pytest.fixture(scope="module")
def env():
    from ocr_rs_b200 import synth
    from ocr_rs_b200.text_detection.model import resnet18
    from oracle import model_oracle as mo
    return synth, resnet18, mo


def _nchw(t):
    return t.numpy() if hasattr(t, "numpy") else t


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("variant", ["tch", "hard_bn"])
def test_small_maps_and_taps(env, mode, variant):
    synth, resnet18, mo = env
    B, H, W = 2, 160, 224  # non-square, several tiles, H/32 and W/32 odd
    w = synth.make_detector_weights(0, variant)
    x = synth.make_noise_images(B, H, W, seed=2)
    x[1] = synth.make_document_images(1, H, W, seed=3, n_boxes=5)[0]
    net = resnet18(w, mode)
    got = net.forward_t(x.reshape(B, 1, H, W))
    ref, taps = mo.detector_forward(w, x.reshape(B, 1, H, W).astype(np.float32), return_taps=True)
    ref = ref.numpy()
    # intermediate taps localise a failure (relative to the tap's own scale)
    for name in ("stem", "x1", "x2", "x3", "x4", "fuse", "bin1"):
        t = taps[name].numpy()
        g = net.tap(name, t.shape)
        rel = np.abs(g - t).max() / max(np.abs(t).max(), 1e-6)
        assert rel < (2e-5 if mode == "fp32" else 3e-2), (name, rel)
    err = np.abs(got - ref).max()
    print(f"{mode}/{variant}: max|dp| = {err:.3e}")
    assert got.shape == ref.shape and err <= TOL[mode], err


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_config1_img55(env, preprocessed, mode):
    """BASELINE config 1: preprocessed_img55.png, random-init weights, single image 800x800."""
    synth, resnet18, mo = env
    img = preprocessed["pre_img55"]
    w = synth.make_detector_weights(0, "tch")
    got = resnet18(w, mode).forward_t(img.reshape(1, 1, 800, 800))
    ref = mo.detector_forward(w, img.reshape(1, 1, 800, 800).astype(np.float32)).numpy()
    err = np.abs(got - ref).max()
    print(f"config1 {mode}: max|dp| = {err:.3e}, mean {np.abs(got - ref).mean():.3e}")
    assert err <= TOL[mode]


def test_f32_input_equals_u8_input(env):
    synth, resnet18, _ = env
    w = synth.make_detector_weights(1, "tch")
    x = synth.make_noise_images(1, 96, 128, seed=5)
    for mode in ("fp32", "bf16"):
        net = resnet18(w, mode)
        a = net.forward_t(x.reshape(1, 1, 96, 128))
        b = net.forward_t(x.reshape(1, 1, 96, 128).astype(np.float32))
        assert (a == b).all()


def test_batch_independence_and_chunking(env):
    """Each image's map must not depend on its batch position (images are sharded by index)."""
    synth, resnet18, _ = env
    w = synth.make_detector_weights(2, "structured1")
    x = synth.make_document_images(5, 96, 96, seed=7, n_boxes=4)
    net = resnet18(w, "bf16")
    full = net.forward_t(x.reshape(5, 1, 96, 96))
    for b in range(5):
        one = net.forward_t(x[b].reshape(1, 1, 96, 96))
        assert (one[0] == full[b]).all()


def test_bad_arguments(env):
    synth, resnet18, _ = env
    from ocr_rs_b200 import OcrbError
    w = synth.make_detector_weights(0, "tch")
    net = resnet18(w, "bf16")
    with pytest.raises(OcrbError):  # FPN adds mis-shape unless H, W are multiples of 32 (model.rs:126-137)
        net.forward_t(np.zeros((1, 1, 100, 100), np.uint8))
    bad = dict(w)
    del bad["layer3.0.downsample.0.weight"]
    with pytest.raises(OcrbError):  # vs.load errors on a missing name
        resnet18(bad, "bf16")
    bad = dict(w)
    bad["in4.weight"] = bad["in4.weight"][:, :128]
    with pytest.raises(OcrbError):
        resnet18(bad, "fp32")


@pytest.mark.parametrize("shape", [(1, 32, 32), (3, 64, 32), (2, 32, 96), (1, 416, 96), (1, 32, 7680)])
def test_tiny_and_odd_shapes(env, shape):
    """Smallest legal maps (one 32x32 cell), shapes whose feature maps are smaller than one MMA tile, and a one-cell-high
    strip whose deepest map is 1 x 240 (the widest tile would not leave room for the weight ring: tests/test_conv_geometry.py)."""
    synth, resnet18, mo = env
    B, H, W = shape
    w = synth.make_detector_weights(3, "hard_bn")
    x = synth.make_noise_images(B, H, W, seed=B * H + W)
    ref = mo.detector_forward(w, x.reshape(B, 1, H, W).astype(np.float32)).numpy()
    for mode in ("fp32", "bf16"):
        got = resnet18(w, mode).forward_t(x.reshape(B, 1, H, W))
        err = np.abs(got - ref).max()
        assert err <= TOL[mode], (mode, shape, err)


def test_config3_batch16(env):
    """BASELINE config 3: batch 16 of 800x800 (8 noise + 8 document images), BF16 mode, map tolerance."""
    synth, resnet18, mo = env
    w = synth.make_detector_weights(0, "structured1")
    x = np.concatenate([synth.make_noise_images(8, 800, 800, seed=2), synth.document_image_shard(0, 8, 800, 800)])
    got = resnet18(w, "bf16").forward_t(x.reshape(16, 1, 800, 800))
    worst = 0.0
    for b0 in range(0, 16, 4):  # oracle in slices to bound host memory
        ref = mo.detector_forward(w, x[b0:b0 + 4].reshape(4, 1, 800, 800).astype(np.float32)).numpy()
        worst = max(worst, float(np.abs(got[b0:b0 + 4] - ref).max()))
    print(f"config3 bf16 batch 16: max|dp| = {worst:.3e}")
    assert worst <= TOL["bf16"]


def test_large_non_square_map(env):
    """A map larger than the benchmark's (1216 x 1600, batch 2): tile / tensor-map arithmetic beyond 800x800."""
    synth, resnet18, mo = env
    B, H, W = 2, 1216, 1600
    w = synth.make_detector_weights(4, "tch")
    x = synth.make_noise_images(B, H, W, seed=9)
    got = resnet18(w, "bf16").forward_t(x.reshape(B, 1, H, W))
    ref = mo.detector_forward(w, x.reshape(B, 1, H, W).astype(np.float32)).numpy()
    err = np.abs(got - ref).max()
    print(f"1216x1600 bf16: max|dp| = {err:.3e}")
    assert err <= TOL["bf16"]


@pytest.mark.parametrize("knobs", [{"OCRB_FUSE_FPN2": "0"}, {"OCRB_FUSE_FPN2": "0", "OCRB_HALO_TS": "0", "OCRB_LATERAL_TS": "0"},
                                   {"OCRB_HALO_CG": "1"}, {"OCRB_CONV": "tc", "OCRB_FUSE_DS": "0"}, {"OCRB_PAIR": "0"},
                                   {"OCRB_STEM": "v2"}, {"OCRB_STEM": "v1"}, {"OCRB_HEAD": "cuda"}, {"OCRB_HEAD": "ss"}])
def test_alternate_kernel_paths(knobs):
    """The library's tuning knobs select older / more literal code paths (the reference's literal FPN graph with the
    lateral kernel, per-thread-store epilogues, single-CTA tiles, the one-box-per-tap engine with separate downsample
    launches, one launch per parity class instead of class pairs, the two older stems: stem_tc2.cu with pixels in the TMEM lanes and the
    im2col stem of stem_tc.cu, the head tail on the CUDA cores / with its second GEMM's operand in shared memory).  They are read once per process, so each combination runs in its own interpreter; same tolerance."""
    import os
    import subprocess
    import sys
    script = r"""
import numpy as np
from ocr_rs_b200 import synth
from ocr_rs_b200.text_detection.model import resnet18
from oracle import model_oracle as mo
worst = 0.0
for (B, H, W), variant in (((2, 160, 224), "hard_bn"), ((1, 416, 96), "tch")):
    w = synth.make_detector_weights(3, variant)
    x = synth.make_noise_images(B, H, W, seed=H + W)
    got = resnet18(w, "bf16").forward_t(x.reshape(B, 1, H, W))
    ref = mo.detector_forward(w, x.reshape(B, 1, H, W).astype(np.float32)).numpy()
    worst = max(worst, float(np.abs(got - ref).max()))
print("WORST", worst)
"""
    env = dict(os.environ, **knobs)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    out = subprocess.run([sys.executable, "-c", script], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    worst = float([l for l in out.stdout.splitlines() if l.startswith("WORST")][-1].split()[1])
    print(knobs, f"max|dp| = {worst:.3e}")
    assert worst <= TOL["bf16"]


@pytest.mark.parametrize("knobs", [{"OCRB_FP32": "cuda"}, {"OCRB_SPLIT_TERMS": "3"}])
def test_fp32_mode_alternate_paths(knobs):
    """FP32 mode runs on the tensor cores by default (operands split into two bf16 terms, conv_tc.cuh); the knobs select the
    CUDA-core fp32 kernels and the three-term (24-bit) split.  Same 1e-4 tolerance, each in its own interpreter."""
    import os
    import subprocess
    import sys
    script = r"""
import numpy as np
from ocr_rs_b200 import synth
from ocr_rs_b200.text_detection.model import resnet18
from oracle import model_oracle as mo
w = synth.make_detector_weights(3, "hard_bn")
x = synth.make_noise_images(2, 160, 224, seed=7)
got = resnet18(w, "fp32").forward_t(x.reshape(2, 1, 160, 224))
ref = mo.detector_forward(w, x.reshape(2, 1, 160, 224).astype(np.float32)).numpy()
print("WORST", float(np.abs(got - ref).max()))
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **knobs)
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    out = subprocess.run([sys.executable, "-c", script], env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    worst = float([l for l in out.stdout.splitlines() if l.startswith("WORST")][-1].split()[1])
    print(knobs, f"max|dp| = {worst:.3e}")
    assert worst <= TOL["fp32"]
