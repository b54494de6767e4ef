"""The decode oracle (oracle/decode_oracle.c, oracle/decode.py) against the reference's own fixtures.

image_ops.rs:805-1008 asserts preprocess_image(text_det/images/*/imgN.jpg, (800, 800)) == test_data/preprocessed_imgN.png
for N in 55, 224, 494, 545.  With jpeg-decoder 0.1.20 restated (stb-style integer IDCT, triangle chroma upsampling on the
component's real size, f32 colour conversion) the whole chain — file bytes -> RGB -> RGBA -> Triangle resize -> luma ->
pad — reproduces all four fixtures bit for bit: three baseline files and one progressive one, 4:2:0 and 4:4:4.
"""
import io

import numpy as np
import pytest

from oracle import decode as dec
from oracle import postproc as pp

ADJUST = {"img55": (800 / 300, 533 / 200), "img224": (600 / 180, 800 / 240), "img494": (800 / 200, 800 / 200),
          "img545": (537 / 184, 800 / 274)}  # image_ops.rs:892-1001


@pytest.fixture(scope="module")
def files(golden_dir):
    z = np.load(golden_dir + "/image_files.npz")
    return {k: z[k] for k in z.files}


def _fixture(files, preprocessed, name):
    return preprocessed["pre_" + name] if "pre_" + name in preprocessed else files["pre_" + name]


@pytest.mark.parametrize("name", ["img55", "img224", "img494", "img545"])
def test_preprocess_fixture_bit_exact_from_file_bytes(files, preprocessed, name):
    rgb = dec.jpeg_decode(files["jpg_" + name].tobytes())
    out, ax, ay = pp.preprocess(dec.to_rgba(rgb), 800, 800)
    assert (ax, ay) == ADJUST[name]
    assert (out == _fixture(files, preprocessed, name)).all()


def test_colour_conversion_variant_is_pinned(files, preprocessed):
    # the fixed-point ycbcr_to_rgb of later jpeg-decoder releases does NOT reproduce the img55 fixture
    rgb = dec.jpeg_decode(files["jpg_img55"].tobytes(), dec.JPEG_COLOR_FIXED)
    out, _, _ = pp.preprocess(dec.to_rgba(rgb), 800, 800)
    assert (out != preprocessed["pre_img55"]).sum() >= 1
    # ... and neither does the normalise-first resize variant
    rgb = dec.jpeg_decode(files["jpg_img55"].tobytes())
    out, _, _ = pp.preprocess(dec.to_rgba(rgb), 800, 800, norm_first=True)
    assert (out != preprocessed["pre_img55"]).sum() >= 1


def test_jpeg_against_libjpeg_within_decoder_noise(files):
    # an independent decoder (Pillow / libjpeg-turbo: different IDCT and upsampling arithmetic) agrees to a few levels
    from PIL import Image
    for k in files:
        if not (k.startswith("jpg_") or k.startswith("synjpg_")):
            continue
        ours = dec.jpeg_decode(files[k].tobytes())
        im = Image.open(io.BytesIO(files[k].tobytes()))
        theirs = np.array(im.convert("RGB" if ours.shape[2] == 3 else "L")).reshape(ours.shape)
        d = np.abs(ours.astype(int) - theirs.astype(int))
        assert ours.shape == theirs.shape
        assert d.max() <= (12 if "synjpg" in k else 4) and d.mean() < 0.8, (k, d.max(), d.mean())


def test_png_decoder(files, preprocessed):
    assert (dec.png_decode(files["png_preprocessed_img55"].tobytes())[..., 0] == preprocessed["pre_img55"]).all()
    g = dec.png_decode(files["png_gt_shrinked_img55"].tobytes())
    assert g.shape == (800, 800, 1) and set(np.unique(g)) <= {0, 255}
    for k in ("rgb", "rgba", "grey", "la", "pal", "pal4", "bilevel", "palt"):
        got = dec.to_rgba(dec.png_decode(files["synpng_" + k].tobytes()))
        assert (got == files["synpng_" + k + "_rgba"]).all(), k


def test_into_luma():
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, size=(5, 7, 3), dtype=np.uint8)
    want = (np.float32(0.2126) * rgb[..., 0].astype(np.float32) + np.float32(0.7152) * rgb[..., 1].astype(np.float32)
            + np.float32(0.0722) * rgb[..., 2].astype(np.float32)).astype(np.uint8)
    assert (dec.to_luma(rgb) == want).all()


def test_malformed_inputs():
    with pytest.raises(ValueError):
        dec.open_image(b"GIF89a....")
    with pytest.raises(ValueError):
        dec.jpeg_decode(b"\xff\xd8\xff\xd9")
