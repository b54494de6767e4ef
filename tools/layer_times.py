#!/usr/bin/env python
"""Per-kernel CUDA-event times of one detector forward (BF16 mode) of N device-resident 800x800 images, printed per
1024 images.   python tools/layer_times.py [N] [name-filter] [bf16|fp32]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ocr_rs_b200 import _ffi, synth  # noqa: E402
from ocr_rs_b200.text_detection.model import resnet18  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flt = sys.argv[2] if len(sys.argv) > 2 else ""
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
ctx = _ffi.default_context(0)
net = resnet18(synth.make_detector_weights(0, "structured"), mode, ctx)
img = torch.from_numpy(synth.document_image_shard(0, n, 800, 800).reshape(n, 1, 800, 800)).cuda()
out = torch.empty((n, 1, 800, 800), dtype=torch.float32, device="cuda")
for _ in range(3):
    net.forward_t(img, out=out)
prof = {}
for _ in range(5):  # per-name minimum over five profiled forwards (single serialised runs show one-off outliers)
    ctx.profile_begin()
    net.forward_t(img, out=out)
    for k, (c, ms) in ctx.profile_end().items():
        if k not in prof or ms < prof[k][1]:
            prof[k] = (c, ms)
tot = 0.0
for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    tot += ms
    if flt in k:
        print(f"{ms * 1024 / n:9.3f} ms/1024  x{c:3d}  {k}")
print(f"{tot * 1024 / n:9.3f} ms/1024  total ({n} images)")
