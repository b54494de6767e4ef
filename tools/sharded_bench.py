#!/usr/bin/env python
"""One process, G GPUs: ocrb_detect_and_read_sharded (the in-library multi-device entry point a single-process
host such as the Rust binary uses) on BASELINE config 4, end to end from pinned host memory.

  python tools/sharded_bench.py [G] [steps]      -> one JSON line

bench.py measures the one-process-per-GPU layout (torchrun); this is the same workload through the C entry point."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from ocr_rs_b200 import _ffi, sharding, synth
    G = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    N, H, W, K = 1024, 800, 800, 4
    imgs = synth.document_image_shard(0, N, H, W)
    L = _ffi.lib()
    # page-locked host buffers from the library's own allocator
    p_img = _ffi.c_p()
    _ffi.check(L.ocrb_host_alloc(imgs.nbytes, C.byref(p_img)))
    himgs = np.ctypeslib.as_array(C.cast(p_img, C.POINTER(C.c_uint8)), shape=imgs.shape)
    himgs[:] = imgs
    adj = np.ones((N, 2))
    sh = sharding.Shards(list(range(G)), synth.make_detector_weights(0, "structured"), synth.make_rec_weights(1), "bf16")
    for _ in range(3):
        res = sh.detect_and_read(himgs, adj, K)
    l0 = sh.launch_count
    t0 = time.perf_counter()
    for _ in range(steps):
        res = sh.detect_and_read(himgs, adj, K)
    dt = time.perf_counter() - t0
    print(json.dumps({"entry": "ocrb_detect_and_read_sharded", "glyphs_per_polygon": K, "n_gpus": G, "steps": steps, "images_per_step": N,
                      "e2e_images_per_s": N * steps / dt, "ms_per_step": 1e3 * dt / steps, "polygons_per_step": int(len(res.all_scores)),
                      "gpu_launches": sh.launch_count - l0, "timing": "host wall clock around the blocking C call (results on the host when it returns)"}))
    sh.close()
    L.ocrb_host_free(p_img)


if __name__ == "__main__":
    main()
