"""CPU: the tiling chosen by the 3x3 convolution kernel (csrc/conv_halo.cu, host side) for every map size a detector
input of up to 4096 x 4096 can produce, and a band of arbitrary sizes: TMA box limits, the sub-tile / operand-stage
arithmetic the kernel relies on, and the shared-memory budget.  No device needed (ocrb_debug_conv_geometry)."""
import ctypes as C
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "ocr_rs_b200", "libocrb.so")):
        g.build()
    from ocr_rs_b200 import _ffi
    return _ffi.lib()


def _sizes():
    s = set()
    for hw in range(32, 4097, 32):          # feature maps of H, W multiples of 32: /4, /8, /16, /32 (+1 column in pair mode)
        for d in (4, 8, 16, 32):
            s.add(hw // d)
    s.update(range(1, 70))                   # small and odd sizes
    return sorted(s)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_geometry_invariants(L, mode):
    G = 4 if mode == 1 else 2
    out = (C.c_int * 10)()
    sizes = _sizes()
    checked = 0
    for ho in sizes:
        for wo in (sizes if ho in (1, 7, 25, 50, 100, 200, 304, 1024) else (ho, max(1, ho // 2), min(1024, 2 * ho))):
            assert L.ocrb_debug_conv_geometry(ho, wo, mode, out) == 0, (ho, wo, mode)
            pw, th, tw, sub_rows, sub_stride, a_stage, a_stages, b_stages, stg, smem = list(out)
            assert tw == pw - 2 and tw >= 1 and th >= 1
            assert pw <= 256 and th + 2 <= 256                    # TMA box dimensions
            if mode == 0:
                assert sub_stride == 128 and th * pw <= G * 128    # linear sub-tiles over the padded pitch
                assert 2 * pw <= 256 and 2 * th <= 256             # the fused downsample reads a stride-2 box of 2 PW x 2 TH pixels
            else:
                assert sub_rows >= 1 and sub_stride == sub_rows * pw <= 128 and th == G * sub_rows
                tile = -(-sub_rows * tw * 128 // 1024) * 1024
                assert stg in (2 * tile * (2 if mode == 2 else 1), 3 * tile * (2 if mode == 2 else 1))
            # the MMA of the last sub-tile, tap (2, 2), reads up to this row of the operand stage
            read_rows = (G - 1) * sub_stride + 128 + 2 * pw + 2
            assert a_stage % 1024 == 0 and a_stage >= max(read_rows, (th + 2) * pw) * 128
            assert a_stages >= 2 and b_stages >= 2
            assert smem <= 227 * 1024 - 4096                       # dynamic limit the launcher requests
            checked += 1
    assert checked > 2000
