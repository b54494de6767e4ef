#!/usr/bin/env python
"""The BASELINE.json configs that bench.py's single line does not cover (it reports config 4):
  1  single image, preprocessed_img55.png, random-init weights       (detector forward, FP32 + BF16)
  2  char_recognition CNN on 4096 glyph crops (28x28)
  3  detection + post-processing, batch 16 of 800x800
  5  post-processing stress, 4096x4096 maps with ~10k components
GPU numbers: CUDA events on the ctx stream, device-resident inputs, 3 warm-ups.  CPU numbers: the
oracle port on this box's host cores (torch threads = all cores; the C post-processing is
single-threaded like the reference's)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ocr_rs_b200 import _ffi, synth  # noqa: E402
from ocr_rs_b200.char_recognition.model import Net  # noqa: E402
from ocr_rs_b200.text_detection.model import resnet18  # noqa: E402
from oracle import model_oracle as mo  # noqa: E402
from oracle import postproc as pp  # noqa: E402


def gpu_time(ctx, fn, reps=10, warm=3):
    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(warm):
        fn()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(reps):
        fn()
    with torch.cuda.stream(stream):
        e1.record()
    ctx.synchronize()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def cpu_time(fn, reps=3):
    fn()
    t = time.time()
    for _ in range(reps):
        fn()
    return (time.time() - t) / reps


def main():
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    ctx = _ffi.default_context(0)
    L = _ffi.lib()
    out = {"host_cores": cores}
    z = np.load(os.path.join(ROOT, "tests", "golden", "preprocessed.npz"))
    img55 = np.ascontiguousarray(z["pre_img55"]).reshape(1, 1, 800, 800)
    w = synth.make_detector_weights(0, "tch")
    # ---- config 1
    d_img = torch.from_numpy(img55).cuda()
    d_out = torch.empty((1, 1, 800, 800), dtype=torch.float32, device="cuda")
    c1 = {}
    for mode in ("fp32", "bf16"):
        net = resnet18(w, mode, ctx)
        t = gpu_time(ctx, lambda: net.forward_t(d_img, out=d_out))
        c1[mode + "_ms"] = t * 1e3
    x55 = torch.from_numpy(img55.astype(np.float32))
    c1["cpu_oracle_ms"] = cpu_time(lambda: mo.detector_forward(w, x55)) * 1e3
    out["config1_single_image_forward"] = c1
    # ---- config 2
    wr = synth.make_rec_weights(1)
    g = synth.make_glyphs(4096, 1, "noise")
    rec = Net(wr, ctx)
    d_g = torch.from_numpy(g).cuda()
    d_am = torch.empty(4096, dtype=torch.int32, device="cuda")
    t = gpu_time(ctx, lambda: _ffi.check(L.ocrb_rec_forward_u8(rec._h, d_g.data_ptr(), 4096, None, d_am.data_ptr(), None)))
    xg = torch.from_numpy(g.astype(np.float32) / np.float32(255))
    tc = cpu_time(lambda: mo.rec_top1(mo.rec_forward(wr, xg)))
    out["config2_rec_4096_glyphs"] = {"gpu_ms": t * 1e3, "gpu_glyphs_per_s": 4096 / t, "cpu_oracle_ms": tc * 1e3, "cpu_glyphs_per_s": 4096 / tc}
    # ---- config 3
    ws = synth.make_detector_weights(0, "structured")
    imgs = synth.document_image_shard(0, 16, 800, 800)
    det = resnet18(ws, "bf16", ctx)
    d_imgs = torch.from_numpy(imgs).cuda()
    adj = np.ones((16, 2))

    def gpu3():
        h = _ffi.c_p()
        _ffi.check(L.ocrb_detect_and_recognize(det._h, None, d_imgs.data_ptr(), _ffi.ptr(adj), 16, 800, 800, None, None, 0, None, C.byref(h)))
        L.ocrb_polygons_free(h)
    t = gpu_time(ctx, gpu3)

    def cpu3():
        pred = mo.detector_forward(ws, torch.from_numpy(imgs[:4].reshape(4, 1, 800, 800)).float()).numpy()
        pp.boxes_and_box_scores(pred, adj[:4])
    tc = cpu_time(cpu3, reps=2) * 4  # 4 of the 16 images timed, scaled
    out["config3_det_postproc_batch16"] = {"gpu_ms": t * 1e3, "gpu_images_per_s": 16 / t, "cpu_oracle_ms_scaled_from_4_images": tc * 1e3, "cpu_images_per_s": 16 / tc}
    # ---- config 5
    prob = np.stack([synth.make_blob_prob_map(4096, 4096, 9000, seed=4, near_thresh=4096, max_w=48, max_h=24)])
    d_prob = torch.from_numpy(prob).cuda()
    adj1 = np.ones((1, 2))
    n_poly = [0]

    def gpu5():
        h = _ffi.c_p()
        _ffi.check(L.ocrb_get_boxes_and_box_scores(ctx.handle, d_prob.data_ptr(), _ffi.ptr(adj1), 1, 4096, 4096, None, C.byref(h)))
        n_poly[0] = int(L.ocrb_polygons_image_offsets(h)[1])
        L.ocrb_polygons_free(h)
    t = gpu_time(ctx, gpu5, reps=5)
    tc = cpu_time(lambda: pp.polygons_from_bitmap(prob[0], pp.binarize(prob[0], 0.6), (1.0, 1.0)), reps=2)
    out["config5_postproc_4096x4096"] = {"polygons": n_poly[0], "gpu_ms_per_map": t * 1e3, "cpu_oracle_ms_per_map_1_thread": tc * 1e3, "speedup": tc / t}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
