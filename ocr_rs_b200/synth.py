"""Seeded synthetic inputs shared by tests, bench.py and the oracle harness.

Pure numpy (no torch, no oracle import): the product path and the checker are
both fed from here so that "same inputs" is literal.

Weight names and shapes are the reference's VarStore names
(/root/reference/src/text_detection/model.rs:65-105,
 /root/reference/src/char_recognition/model.rs:12-25).
Initialisation follows tch 0.3.0 defaults as recalled in SURVEY.md §8(d):
conv / conv-transpose / linear weights U(-b, b) with b = sqrt(1 / fan_in),
fan_in = prod(shape[1:]); conv biases 0; linear bias U(+-1/sqrt(in));
batch-norm weight U(0,1), bias 0, running_mean 0, running_var 1.
"""
from __future__ import annotations

import numpy as np

VALUES = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789"  # utils.rs:7

# --------------------------------------------------------------------------- detector


def detector_weight_specs():
    """[(name, shape)] in VarStore creation order (model.rs:65-105)."""
    specs = [("conv1.weight", (64, 1, 7, 7))]
    specs += _bn("bn1", 64)
    cin = 64
    for li, c in zip((1, 2, 3, 4), (64, 128, 256, 512)):
        for b in (0, 1):
            p = f"layer{li}.{b}"
            bc_in = cin if b == 0 else c
            specs.append((f"{p}.conv1.weight", (c, bc_in, 3, 3)))
            specs += _bn(f"{p}.bn1", c)
            specs.append((f"{p}.conv2.weight", (c, c, 3, 3)))
            specs += _bn(f"{p}.bn2", c)
            if b == 0 and (li != 1):
                specs.append((f"{p}.downsample.0.weight", (c, bc_in, 1, 1)))
                specs += _bn(f"{p}.downsample.1", c)
        cin = c
    specs += [
        ("in5.weight", (256, 512, 1, 1)),
        ("in4.weight", (256, 256, 1, 1)),
        ("in3.weight", (256, 128, 1, 1)),
        ("in2.weight", (256, 64, 1, 1)),
        ("out5.weight", (64, 256, 3, 3)),
        ("out4.weight", (64, 256, 3, 3)),
        ("out3.weight", (64, 256, 3, 3)),
        ("out2.weight", (64, 256, 3, 3)),
        ("bin_conv1.weight", (64, 256, 3, 3)),
    ]
    specs += _bn("bin_bn1", 64)
    specs += [("bin_conv_tr1.weight", (64, 64, 2, 2)), ("bin_conv_tr1.bias", (64,))]
    specs += _bn("bin_bn2", 64)
    specs += [("bin_conv_tr2.weight", (64, 1, 2, 2)), ("bin_conv_tr2.bias", (1,))]
    return specs


def _bn(prefix, c):
    return [
        (f"{prefix}.weight", (c,)),
        (f"{prefix}.bias", (c,)),
        (f"{prefix}.running_mean", (c,)),
        (f"{prefix}.running_var", (c,)),
    ]


def _uniform_fan_in(rng, shape):
    fan_in = int(np.prod(shape[1:]))
    b = np.sqrt(1.0 / fan_in)
    return rng.uniform(-b, b, size=shape).astype(np.float32)


def make_detector_weights(seed=0, variant="tch"):
    """dict name -> float32 array.

    variant:
      "tch"        plain tch-style random init
      "hard_bn"    + running_mean ~ N(0,1)*0.1, running_var ~ U(0.5,2), bias ~ N(0,0.1)
                   (exercises batch-norm folding)
      "structured" tied 2x2 transposed-conv taps and bin_conv_tr2 gain 64
                   (SURVEY.md §8(d) cfg 3: blocky maps with surviving polygons)
      "structured1" tied taps, gain 1 (map-tolerance checks)
    """
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in detector_weight_specs():
        leaf = name.rsplit(".", 1)[1]
        is_bn = len(shape) == 1 and not name.startswith("bin_conv_tr")
        if is_bn:
            if leaf == "weight":
                w[name] = rng.uniform(0.0, 1.0, size=shape).astype(np.float32)
            elif leaf == "running_var":
                w[name] = np.ones(shape, np.float32)
            else:
                w[name] = np.zeros(shape, np.float32)
        elif leaf == "bias":
            w[name] = np.zeros(shape, np.float32)
        else:
            w[name] = _uniform_fan_in(rng, shape)
    if variant == "hard_bn":
        rng2 = np.random.default_rng(seed + 1000)
        for name in list(w):
            if name.endswith("running_mean"):
                w[name] = (0.1 * rng2.standard_normal(w[name].shape)).astype(np.float32)
            elif name.endswith("running_var"):
                w[name] = rng2.uniform(0.5, 2.0, size=w[name].shape).astype(np.float32)
            elif name.endswith(".bias") and not name.startswith("bin_conv_tr"):
                w[name] = (0.1 * rng2.standard_normal(w[name].shape)).astype(np.float32)
        w["bin_conv_tr1.bias"] = (0.05 * rng2.standard_normal(64)).astype(np.float32)
        w["bin_conv_tr2.bias"] = (0.05 * rng2.standard_normal(1)).astype(np.float32)
    elif variant in ("structured", "structured1"):
        for n in ("bin_conv_tr1.weight", "bin_conv_tr2.weight"):
            w[n] = np.ascontiguousarray(np.broadcast_to(w[n][:, :, :1, :1], w[n].shape)).copy()
        if variant == "structured":
            w["bin_conv_tr2.weight"] = (w["bin_conv_tr2.weight"] * 64.0).astype(np.float32)
    elif variant != "tch":
        raise ValueError(variant)
    return w


# --------------------------------------------------------------------------- char-rec

REC_CANONICAL = [
    ("conv1.bias", (32,)),
    ("conv1.weight", (32, 1, 5, 5)),
    ("conv2.bias", (64,)),
    ("conv2.weight", (64, 32, 5, 5)),
    ("fc1.bias", (512,)),
    ("fc1.weight", (512, 1024)),
    ("fc2.bias", (62,)),
    ("fc2.weight", (62, 512)),
]
# tch de-duplicated names when all four layers share one path (SURVEY.md Appendix B,
# char_recognition/model.rs:14-17).  The "__<n>" numbering depends on the creation order inside nn::conv /
# nn::linear, which differs between tch versions: ocrb_rec_create ignores the suffix and places weight* / bias*
# tensors by element count, so both orders (and any other numbering) load.
REC_VARSTORE_ALIASES = [  # bias before weight in every layer
    "bias", "weight", "bias__2", "weight__3", "bias__4", "weight__5", "bias__6", "weight__7",
]
REC_VARSTORE_ALIASES_WEIGHT_FIRST_LINEAR = [  # nn::conv bias-first, nn::linear weight-first (tch 0.3.0 as read by the advisor)
    "bias", "weight", "bias__2", "weight__3", "bias__5", "weight__4", "bias__7", "weight__6",
]


def make_rec_weights(seed=1):
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in REC_CANONICAL:
        if name.endswith("weight"):
            w[name] = _uniform_fan_in(rng, shape)
        elif name.startswith("fc"):
            fan_in = 1024 if name.startswith("fc1") else 512
            b = 1.0 / np.sqrt(fan_in)
            w[name] = rng.uniform(-b, b, size=shape).astype(np.float32)
        else:
            w[name] = np.zeros(shape, np.float32)
    return w


def make_glyphs(n, seed=1, kind="noise"):
    """[n, 784] uint8 glyph crops (28x28). kind: noise | strokes."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, size=(n, 784), dtype=np.uint8)
    g = np.zeros((n, 28, 28), np.uint8)
    yy, xx = np.mgrid[0:28, 0:28]
    for i in range(n):
        for _ in range(int(rng.integers(2, 5))):
            x0, y0, x1, y1 = rng.uniform(3, 25, size=4)
            t = rng.uniform(1.0, 2.5)
            dx, dy = x1 - x0, y1 - y0
            l2 = dx * dx + dy * dy + 1e-6
            u = np.clip(((xx - x0) * dx + (yy - y0) * dy) / l2, 0, 1)
            d = np.hypot(xx - (x0 + u * dx), yy - (y0 + u * dy))
            g[i] = np.maximum(g[i], (255 * np.clip(t - d, 0, 1)).astype(np.uint8))
    return g.reshape(n, 784)


# --------------------------------------------------------------------------- images


def make_noise_images(n, h=800, w=800, seed=2):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, h, w), dtype=np.uint8)


def make_document_images(n, h=800, w=800, seed=3, n_boxes=30):
    """Dark background + bright rotated rectangles / strokes (cfg 3/4 'document' images)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for i in range(n):
        img = rng.integers(0, 40, size=(h, w)).astype(np.float32)
        for _ in range(n_boxes):
            cx, cy = rng.uniform(40, w - 40), rng.uniform(40, h - 40)
            bw, bh = rng.uniform(30, 160), rng.uniform(8, 40)
            a = rng.uniform(-0.5, 0.5)
            ca, sa = np.cos(a), np.sin(a)
            u = (xx - cx) * ca + (yy - cy) * sa
            v = -(xx - cx) * sa + (yy - cy) * ca
            m = (np.abs(u) < bw / 2) & (np.abs(v) < bh / 2)
            img[m] = rng.uniform(150, 255)
        out[i] = np.clip(img, 0, 255).astype(np.uint8)
    return out


def make_document_images_fast(n, h=800, w=800, seed=3, n_boxes=30):
    """Same kind of image as make_document_images, drawn box-window by box-window (about
    50x faster); used where hundreds of images are needed (bench.py, cfg 4)."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        img = rng.integers(0, 40, size=(h, w), dtype=np.uint8)
        for _ in range(n_boxes):
            cx, cy = rng.uniform(40, w - 40), rng.uniform(40, h - 40)
            bw, bh = rng.uniform(30, 160), rng.uniform(8, 40)
            a = rng.uniform(-0.5, 0.5)
            r = int(np.ceil(np.hypot(bw, bh) / 2)) + 1
            x0, x1 = max(0, int(cx) - r), min(w, int(cx) + r + 1)
            y0, y1 = max(0, int(cy) - r), min(h, int(cy) + r + 1)
            yy, xx = np.mgrid[y0:y1, x0:x1].astype(np.float32)
            ca, sa = np.cos(a), np.sin(a)
            u = (xx - cx) * ca + (yy - cy) * sa
            v = -(xx - cx) * sa + (yy - cy) * ca
            m = (np.abs(u) < bw / 2) & (np.abs(v) < bh / 2)
            img[y0:y1, x0:x1][m] = np.uint8(rng.uniform(150, 255))
        out[i] = img
    return out


def document_image_shard(first, count, h=800, w=800, seed=3, unique=128):
    """Images [first, first+count) of the endless synthetic 'document' set: image i is base
    image i % unique rolled by a shift that depends on i // unique, so it is a function of the
    global index only (identical no matter how the index range is sharded across GPUs)."""
    base = make_document_images_fast(min(unique, first + count), h, w, seed)  # sequential rng: a prefix is stable
    out = np.empty((count, h, w), np.uint8)
    for k in range(count):
        i = first + k
        rep = i // unique
        img = base[i % unique]
        out[k] = img if rep == 0 else np.roll(img, (37 * rep % h, 53 * rep % w), (0, 1))
    return out


def make_blob_prob_map(h=800, w=800, n_blobs=40, seed=4, frame=1, ring_frac=0.1, near_thresh=64, max_w=120, max_h=60):
    """Synthetic probability map (cfg 5 generator).

    Background U(0,0.5); blobs (rotated rectangles / ellipses / rings) with interior
    U(0.62,0.99); `near_thresh` pixels set within 1e-6 of 0.6; a `frame`-px border is
    kept at background so no component touches the image edge.
    """
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.0, 0.5, size=(h, w)).astype(np.float32)
    occ = np.zeros((h, w), bool)
    placed = 0
    tries = 0
    while placed < n_blobs and tries < n_blobs * 30:
        tries += 1
        bw, bh = rng.uniform(8, max_w), rng.uniform(8, max_h)
        r = int(np.ceil(np.hypot(bw, bh) / 2)) + 3
        cx = int(rng.integers(r + frame, w - r - frame))
        cy = int(rng.integers(r + frame, h - r - frame))
        ys, xs = slice(cy - r, cy + r + 1), slice(cx - r, cx + r + 1)
        if occ[ys, xs].any():
            continue
        yy, xx = np.mgrid[-r:r + 1, -r:r + 1].astype(np.float32)
        a = rng.uniform(0, np.pi)
        ca, sa = np.cos(a), np.sin(a)
        u = xx * ca + yy * sa
        v = -xx * sa + yy * ca
        kind = rng.uniform()
        if kind < 0.5:
            m = (np.abs(u) <= bw / 2) & (np.abs(v) <= bh / 2)
        else:
            q = (u / (bw / 2)) ** 2 + (v / (bh / 2)) ** 2
            m = q <= 1.0
            if kind > 1.0 - ring_frac and min(bw, bh) > 24:
                m &= q >= 0.35
        if m.sum() < 12:
            continue
        vals = rng.uniform(0.62, 0.99, size=m.shape).astype(np.float32)
        sub = p[ys, xs]
        sub[m] = vals[m]
        occ[ys, xs] = True
        placed += 1
    if near_thresh:
        ys = rng.integers(frame, h - frame, size=near_thresh)
        xs = rng.integers(frame, w - frame, size=near_thresh)
        t = np.float32(0.6)
        choices = np.array([t, np.nextafter(t, np.float32(1)), np.nextafter(t, np.float32(0)),
                            np.float32(0.6000005), np.float32(0.5999995)], np.float32)
        free = ~occ[ys, xs]
        p[ys[free], xs[free]] = choices[rng.integers(0, len(choices), size=int(free.sum()))]
    return p


def make_random_bitmap(h, w, seed, density=0.5, smooth=0):
    """Random binary image for contour / CCL fuzzing (touches the frame on purpose)."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(size=(h, w))
    for _ in range(smooth):
        a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1)) / 5
    if smooth:
        thr = np.quantile(a, 1 - density)
        return (a > thr).astype(np.uint8)
    return (a < density).astype(np.uint8)
