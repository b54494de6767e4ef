"""Regenerates tests/golden/*.npz from the reference's own test fixtures.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box).  The arrays are the reference's test DATA (test_data/*.png, *.jpg), decoded and stored
as numpy so that tests need no image decoder and no /root/reference at run time:

  gt_shrinked_img55.npz   bitmap used by metrics.rs:510-646 (golden polygons + scores)
  preprocessed.npz        preprocessed_img{55,545}.png (image_ops.rs:805-1008 expectations,
                          and BASELINE config 1 input) + the decoded source JPEGs (PIL/libjpeg)
  gt_shrinked_others.npz  gt_shrinked_img{224,494,545}.png: extra real blob bitmaps
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


if __name__ == "__main__":
    main()
