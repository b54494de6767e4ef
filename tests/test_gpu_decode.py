"""File decode on the device (csrc/decode.cu) through the C ABI against the oracle and the reference's fixtures:
image::open(file)?.into_rgba() / .into_luma() and preprocess_image(file, dims) (image_ops.rs:73-85, 188-220)."""
import numpy as np
import pytest

from ocr_rs_b200 import image_ops
from oracle import decode as dec
from oracle import postproc as pp

pytestmark = pytest.mark.gpu

ADJUST = {"img55": (800 / 300, 533 / 200), "img224": (600 / 180, 800 / 240), "img494": (800 / 200, 800 / 200),
          "img545": (537 / 184, 800 / 274)}  # image_ops.rs:892-1001


def _encoded(image_files):
    return {k: v.tobytes() for k, v in image_files.items()
            if k.split("_")[0] in ("jpg", "synjpg", "png", "synpng") and not (k.endswith("_rgba") and k[:-5] in image_files)}


def test_decode_rgba_and_luma_bit_exact(image_files):
    enc = _encoded(image_files)
    names = sorted(enc)
    got = image_ops.decode_images([enc[k] for k in names], "rgba")
    gotl = image_ops.decode_images([enc[k] for k in names], "luma")
    for k, a, l in zip(names, got, gotl):
        px = dec.open_image(enc[k])
        assert a.shape == px.shape[:2] + (4,), k
        assert (a == dec.to_rgba(px)).all(), k
        assert (l == dec.to_luma(px)).all(), k


def test_preprocess_files_reproduces_the_reference_fixtures(image_files, preprocessed):
    # image_ops.rs:805-1008: preprocess_image(imgN.jpg, (800, 800)) == preprocessed_imgN.png — from the file BYTES, in one batch
    names = ["img55", "img224", "img494", "img545"]
    out, adj = image_ops.preprocess_files([image_files["jpg_" + n].tobytes() for n in names], (800, 800))
    for i, n in enumerate(names):
        want = preprocessed["pre_" + n] if "pre_" + n in preprocessed else image_files["pre_" + n]
        assert (out[i] == want).all(), n
        assert tuple(adj[i]) == ADJUST[n]
    # single-image form with the reference's signature
    img, ax, ay = image_ops.preprocess_image(image_files["jpg_img55"].tobytes(), (800, 800))
    assert (img == preprocessed["pre_img55"]).all() and (ax, ay) == ADJUST["img55"]


def test_preprocess_files_mixed_batch_and_path(image_files, tmp_path):
    enc = _encoded(image_files)
    names = ["synjpg_420_odd", "synpng_rgba", "synjpg_grey", "png_preprocessed_img55", "synjpg_422_progressive_restart", "synpng_pal4",
             "synjpg_tiny", "synjpg_420_restart"]
    out, adj = image_ops.preprocess_files([enc[k] for k in names], (320, 256))
    for i, k in enumerate(names):
        want, ax, ay = pp.preprocess(dec.to_rgba(dec.open_image(enc[k])), 320, 256)
        assert (out[i] == want).all(), k
        assert (adj[i, 0], adj[i, 1]) == (ax, ay)
    p = tmp_path / "a.jpg"
    p.write_bytes(enc["synjpg_422"])
    img, ax, ay = image_ops.preprocess_image(str(p), (96, 64))
    want, wx, wy = pp.preprocess(dec.to_rgba(dec.open_image(enc["synjpg_422"])), 96, 64)
    assert (img == want).all() and (ax, ay) == (wx, wy)


def test_load_image_as_tensor_from_file(image_files, tmp_path):
    data = image_files["synjpg_444_q50"].tobytes()
    t = image_ops.load_image_as_tensor(data)
    luma = dec.to_luma(dec.open_image(data))
    assert t.shape == (1, luma.size)
    assert (t[0] == luma.reshape(-1).astype(np.float32) / np.float32(255.0)).all()
    with pytest.raises(FileNotFoundError):
        image_ops.load_image_as_tensor(str(tmp_path / "missing.png"))


def test_unsupported_file_is_an_error():
    from ocr_rs_b200 import _ffi
    with pytest.raises(_ffi.OcrbError):
        image_ops.preprocess_files([b"GIF89a" + b"\0" * 64], (64, 64))
