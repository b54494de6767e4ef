"""Regenerates tests/golden/*.npz from the reference's own test fixtures.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box).  The arrays are the reference's test DATA (test_data/*.png, *.jpg), decoded and stored
as numpy so that tests need no image decoder and no /root/reference at run time:

  gt_shrinked_img55.npz   bitmap used by metrics.rs:510-646 (golden polygons + scores)
  preprocessed.npz        preprocessed_img{55,545}.png (image_ops.rs:805-1008 expectations,
                          and BASELINE config 1 input) + the decoded source JPEGs (PIL/libjpeg)
  gt_shrinked_others.npz  gt_shrinked_img{224,494,545}.png: extra real blob bitmaps
  text_det_gts.npz        the four ground-truth polygon files test_data/text_det/gts/*/*.txt, the
                          adjust values image_ops.rs:892-1001 asserts for them, and mask_img*.png:
                          inputs / expectations of generate_gt_and_mask_images (image_ops.rs:222-277),
                          whose outputs gt_shrinked_*.png / mask_*.png the reference pins at :805-1008
  image_files.npz         the reference's four source JPEGs and two of its PNGs as RAW FILE BYTES (inputs of the decoders:
                          image::open at image_ops.rs:193 / :78), the two preprocessed fixtures preprocessed.npz lacks, and
                          small synthetic JPEG / PNG files written by Pillow that cover what the reference's files do not
                          (4:2:2, grey, restart intervals, odd sizes; PNG colour types, bit depths and filters) together
                          with Pillow's decode of the PNGs
"""
import io

import numpy as np
from PIL import Image

REF = "/root/reference/test_data"


def gray(name):
    return np.array(Image.open(f"{REF}/{name}.png").convert("L"))


def main():
    g = gray("gt_shrinked_img55")
    np.savez_compressed("tests/golden/gt_shrinked_img55.npz", bits=np.packbits(g > 0), shape=np.array(g.shape))
    others = {}
    for n in ("img224", "img494", "img545"):
        a = gray(f"gt_shrinked_{n}")
        others[n] = np.packbits(a > 0)
        others[n + "_shape"] = np.array(a.shape)
    np.savez_compressed("tests/golden/gt_shrinked_others.npz", **others)
    pre = {}
    for n, sub in (("img55", "train"), ("img545", "test")):
        pre["pre_" + n] = gray(f"preprocessed_{n}")
        pre["src_" + n] = np.array(Image.open(f"{REF}/text_det/images/{sub}/{n}.jpg").convert("RGBA"))
    np.savez_compressed("tests/golden/preprocessed.npz", **pre)
    files = {}
    for n, sub in (("img55", "train"), ("img224", "train"), ("img494", "test"), ("img545", "test")):
        files["jpg_" + n] = np.frombuffer(open(f"{REF}/text_det/images/{sub}/{n}.jpg", "rb").read(), np.uint8)
    for n in ("preprocessed_img55", "gt_shrinked_img55"):
        files["png_" + n] = np.frombuffer(open(f"{REF}/{n}.png", "rb").read(), np.uint8)
    for n in ("img224", "img494"):
        files["pre_" + n] = gray(f"preprocessed_{n}")
    rng = np.random.default_rng(7)

    def smooth(h, w, c):
        a = rng.integers(0, 256, size=(h // 4 + 2, w // 4 + 2, c)).astype(np.float32)
        a = np.kron(a, np.ones((4, 4, 1), np.float32))[:h, :w]
        a += rng.normal(0, 6, size=a.shape)
        return np.clip(a, 0, 255).astype(np.uint8)

    def save(img, fmt, **kw):
        b = io.BytesIO()
        img.save(b, fmt, **kw)
        return np.frombuffer(b.getvalue(), np.uint8)

    rgb = smooth(61, 83, 3)
    files["synjpg_420_odd"] = save(Image.fromarray(rgb), "JPEG", quality=85, subsampling=2)
    files["synjpg_422"] = save(Image.fromarray(rgb), "JPEG", quality=90, subsampling=1)
    files["synjpg_444_q50"] = save(Image.fromarray(rgb), "JPEG", quality=50, subsampling=0)
    files["synjpg_grey"] = save(Image.fromarray(rgb[..., 0]), "JPEG", quality=80)
    files["synjpg_420_restart"] = save(Image.fromarray(smooth(120, 200, 3)), "JPEG", quality=75, subsampling=2, restart_marker_blocks=3)
    files["synjpg_420_progressive"] = save(Image.fromarray(smooth(97, 131, 3)), "JPEG", quality=92, subsampling=2, progressive=True)
    files["synjpg_422_progressive_restart"] = save(Image.fromarray(smooth(64, 48, 3)), "JPEG", quality=60, subsampling=1, progressive=True,
                                                  restart_marker_blocks=2)
    files["synjpg_tiny"] = save(Image.fromarray(smooth(8, 8, 3)[:1, :1]), "JPEG", quality=90, subsampling=2)
    pngs = {"rgb": Image.fromarray(rgb), "rgba": Image.fromarray(smooth(37, 23, 4)), "grey": Image.fromarray(rgb[..., 1]),
            "la": Image.fromarray(smooth(19, 31, 2), "LA"), "pal": Image.fromarray(rgb).quantize(64),
            "pal4": Image.fromarray(rgb).quantize(13), "bilevel": Image.fromarray(rgb[..., 0] > 128)}
    for k, im in pngs.items():
        files["synpng_" + k] = save(im, "PNG")
        files["synpng_" + k + "_rgba"] = np.array(im.convert("RGBA"))
    palt = Image.fromarray(rgb).quantize(32)
    files["synpng_palt"] = save(palt, "PNG", transparency=bytes(range(0, 256, 8)))
    files["synpng_palt_rgba"] = np.array(Image.open(io.BytesIO(files["synpng_palt"].tobytes())).convert("RGBA"))
    np.savez_compressed("tests/golden/image_files.npz", **files)
    gts = {}
    dims = {"img224": ("train", (180, 240), (600, 800)), "img55": ("train", (300, 200), (800, 533)),
            "img494": ("test", (200, 200), (800, 800)), "img545": ("test", (184, 274), (537, 800))}
    for n, (sub, orig, resized) in dims.items():
        rows = []
        for row in open(f"{REF}/text_det/gts/{sub}/{n}.jpg.txt").read().split("\n"):
            if not row:
                continue
            vals = []
            for v in row.split(",")[:-1]:  # image_ops.rs:292-296: flat_map(parse) drops what does not parse
                try:
                    vals.append(int(v))
                except ValueError:
                    pass
            rows.append(np.array(vals[: len(vals) // 2 * 2], np.int32).reshape(-1, 2))
        gts[n + "_counts"] = np.array([len(r) for r in rows], np.int32)
        gts[n + "_points"] = np.concatenate(rows)
        gts[n + "_orig"] = np.array(orig, np.float64)
        gts[n + "_resized"] = np.array(resized, np.float64)
        m = gray(f"mask_{n}")
        gts[n + "_mask_bits"] = np.packbits(m > 0)
        assert set(np.unique(m)) <= {0, 255} and set(np.unique(gray(f"gt_shrinked_{n}"))) <= {0, 255}
    np.savez_compressed("tests/golden/text_det_gts.npz", **gts)


if __name__ == "__main__":
    main()
