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
"""
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
