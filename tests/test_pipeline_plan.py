"""Host logic of the batched entry points (ocrb_detect_and_read / ocrb_detect_and_recognize; the reference's batched loop is
text_detection/mod.rs:188-204): how a batch is cut into post-processing groups and forward chunks.  No device needed
(ocrb_debug_pipeline_plan).  The plans are checked for what every plan must satisfy — each image forwarded exactly once,
chunks never across a group, sizes within the kernels' limits — and for the default schedule bench.py's numbers were taken
with (per-rank batches of 1024 / 512 / 256 / 128 images at 1 / 2 / 4 / 8 GPUs)."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L():
    from ocr_rs_b200 import _ffi
    return _ffi.lib()


def plan(L, B, H=800, W=800, bf16=1, host=1):
    group, n = C.c_int(), C.c_int()
    chunks = np.zeros(4096, np.int32)
    rc = L.ocrb_debug_pipeline_plan(B, H, W, bf16, host, C.byref(group), chunks.ctypes.data_as(C.c_void_p), len(chunks), C.byref(n))
    assert rc == 0
    return group.value, [int(c) for c in chunks[: n.value]]


def test_default_schedule(L):
    # one post-processing group per <= 256 images; host images ramp 16, 48, 144 and then run at the chunk size
    assert plan(L, 1024) == (256, [16, 48, 144, 48, 256, 256, 256])
    assert plan(L, 512) == (256, [16, 48, 144, 48, 256])
    assert plan(L, 256) == (256, [16, 48, 144, 48])
    assert plan(L, 128) == (128, [16, 48, 64])
    # device-resident images: nothing to hide, whole chunks
    assert plan(L, 1024, host=0) == (256, [256, 256, 256, 256])
    assert plan(L, 128, host=0) == (128, [128])
    assert plan(L, 300, host=0) == (256, [256, 44])
    # FP32 mode forwards 16 images at a time (activation memory): no ramp below the chunk size
    assert plan(L, 40, bf16=0) == (40, [16, 16, 8])
    assert plan(L, 1) == (1, [1])


@pytest.mark.parametrize("B", [1, 2, 15, 16, 17, 63, 64, 65, 127, 128, 200, 255, 256, 257, 511, 512, 513, 1000, 1024, 3000])
@pytest.mark.parametrize("hw", [(160, 160), (800, 800), (4096, 4096)])
@pytest.mark.parametrize("host", [0, 1])
def test_plan_invariants(L, B, hw, host):
    H, W = hw
    for bf16 in (0, 1):
        group, chunks = plan(L, B, H, W, bf16, host)
        assert sum(chunks) == B and all(c > 0 for c in chunks)
        assert 1 <= group <= 256 and group * H * W < 2 ** 31  # post-processing index arithmetic is 32-bit per group
        assert max(chunks) <= (256 if bf16 else 16)
        # chunks never straddle a group boundary
        pos = 0
        for c in chunks:
            assert pos // group == (pos + c - 1) // group
            pos += c


def test_capacity_error(L):
    group, n = C.c_int(), C.c_int()
    chunks = np.zeros(2, np.int32)
    assert L.ocrb_debug_pipeline_plan(1024, 800, 800, 1, 1, C.byref(group), chunks.ctypes.data_as(C.c_void_p), 2, C.byref(n)) == -3
    assert n.value == 7
