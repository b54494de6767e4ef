#!/usr/bin/env python
"""Quick check of the glyph net on the GPU: accuracy vs the torch restatement and throughput, for the tensor-core
(default) and the fp32 CUDA-core path (OCRB_REC=fp32).   python tools/rec_check.py [n_glyphs]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from ocr_rs_b200 import synth
    from ocr_rs_b200.char_recognition.model import Net
    from oracle import model_oracle as mo
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    w = synth.make_rec_weights(1)
    g = synth.make_glyphs(n, 1, "noise")
    net = Net(w)
    logits, argmax, prob = net.predict(g)
    m = min(n, 8192)
    ref = mo.rec_forward(w, g[:m].astype(np.float32) / np.float32(255.0)).numpy()
    print(f"path={os.environ.get('OCRB_REC', 'tc')} n={n}: max|dlogit| = {np.abs(logits[:m] - ref).max():.3e} (|logit| max {np.abs(ref).max():.3f}), "
          f"argmax mismatches {(argmax[:m] != ref.argmax(-1)).sum()}")
    dg = torch.from_numpy(g).cuda()
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    from ocr_rs_b200 import _ffi
    L = _ffi.lib()
    for _ in range(3):
        _ffi.check(L.ocrb_rec_forward_u8(net._h, _ffi.ptr(dg), n, None, _ffi.ptr(out), None))
    t0 = time.perf_counter()
    for _ in range(10):
        _ffi.check(L.ocrb_rec_forward_u8(net._h, _ffi.ptr(dg), n, None, _ffi.ptr(out), None))
    dt = (time.perf_counter() - t0) / 10
    print(f"  {1e3 * dt:.3f} ms per {n} glyphs = {n / dt / 1e6:.2f} M glyphs/s")
    net.ctx.profile_begin()
    _ffi.check(L.ocrb_rec_forward_u8(net._h, _ffi.ptr(dg), n, None, _ffi.ptr(out), None))
    for k, (c, ms) in net.ctx.profile_end().items():
        print(f"    {ms:8.3f} ms x{c} {k}")


if __name__ == "__main__":
    main()
