#!/usr/bin/env python
"""Two detector forwards (BF16 mode) of N device-resident 800x800 document images — the command
that is run under ncu for profiles/*_ncu_forward_*: the first forward warms up (weight prep,
tensor maps), the second is the one captured (--launch-skip 27 --launch-count 27 with
-k regex:stem_tc|conv_tc|conv_halo|conv_lateral)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ocr_rs_b200 import _ffi, synth  # noqa: E402
from ocr_rs_b200.text_detection.model import resnet18  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = _ffi.default_context(0)
w = synth.make_detector_weights(0, "structured1")
net = resnet18(w, "bf16", ctx)
img = torch.from_numpy(synth.document_image_shard(0, n, 800, 800).reshape(n, 1, 800, 800)).cuda()
out = torch.empty((n, 1, 800, 800), dtype=torch.float32, device="cuda")
for _ in range(2):
    net.forward_t(img, out=out)
    ctx.synchronize()
print("forward ok", float(out.mean()))
