#!/usr/bin/env python
"""FP32-accuracy mode of the detector: map error against the torch-CPU restatement and throughput, for the tensor-core
split-operand path (default; OCRB_SPLIT_TERMS=2|3) and the CUDA-core path (OCRB_FP32=cuda).
  python tools/fp32_mode_check.py [n_images]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocr_rs_b200 import _ffi, synth  # noqa: E402
from ocr_rs_b200.text_detection.model import resnet18  # noqa: E402
from oracle import model_oracle as mo  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ctx = _ffi.default_context(0)
label = f"OCRB_FP32={os.environ.get('OCRB_FP32', 'tc')} terms={os.environ.get('OCRB_SPLIT_TERMS', '2')}"
for variant in ("tch", "hard_bn", "structured1"):
    w = synth.make_detector_weights(0, variant)
    net = resnet18(w, "fp32", ctx)
    x = synth.make_document_images(2, 800, 800, seed=5).reshape(2, 1, 800, 800)
    got = net.forward_t(x)
    ref = mo.detector_forward(w, x.astype(np.float32)).numpy()
    print(f"{label} {variant}: max|dp| = {np.abs(got - ref).max():.3e}")
net = resnet18(synth.make_detector_weights(0, "structured"), "fp32", ctx)
img = torch.from_numpy(synth.document_image_shard(0, n, 800, 800).reshape(n, 1, 800, 800)).cuda()
out = torch.empty((n, 1, 800, 800), dtype=torch.float32, device="cuda")
for _ in range(2):
    net.forward_t(img, out=out)
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    net.forward_t(img, out=out)
ctx.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"{label}: {n / dt:.1f} images/s ({1e3 * dt / n:.3f} ms per 800x800 image)")
