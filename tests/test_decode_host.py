"""Host stage of the file decode in libocrb.so (csrc/decode.cu: JPEG entropy decoding, PNG inflate + filters) against
the oracle — no GPU needed: ocrb_debug_decode_host returns what the host hands to the device kernels."""
import ctypes as C
import io
import zlib

import numpy as np
import pytest

from ocr_rs_b200 import _ffi, image_ops
from oracle import decode as dec


def host_stage(data: bytes, dtype):
    L = _ffi.lib()
    buf = np.frombuffer(data, np.uint8)
    need = C.c_size_t()
    _ffi.check(L.ocrb_debug_decode_host(buf.ctypes.data, len(data), None, 0, C.byref(need)))
    out = np.empty(need.value // np.dtype(dtype).itemsize, dtype)
    _ffi.check(L.ocrb_debug_decode_host(buf.ctypes.data, len(data), out.ctypes.data, need.value, C.byref(need)))
    return out


def test_jpeg_coefficients_equal_the_oracle(image_files):
    n = 0
    for k, v in image_files.items():
        if k.startswith("jpg_") or k.startswith("synjpg_"):
            got = host_stage(v.tobytes(), np.int16)
            want = dec.jpeg_coefficients(v.tobytes())
            assert got.shape == want.shape and (got == want).all(), k
            n += 1
    assert n >= 12


def test_png_pixels_equal_the_oracle_and_pillow(image_files, preprocessed):
    got = host_stage(image_files["png_preprocessed_img55"].tobytes(), np.uint8)
    assert (got.reshape(800, 800) == preprocessed["pre_img55"]).all()
    for k in ("rgb", "rgba", "grey", "la", "pal", "pal4", "bilevel", "palt"):
        data = image_files["synpng_" + k].tobytes()
        want = dec.png_decode(data)
        got = host_stage(data, np.uint8).reshape(want.shape)
        assert (got == want).all(), k
        assert (dec.to_rgba(got) == image_files["synpng_" + k + "_rgba"]).all(), k
        w, h, c = image_ops.image_info(data)
        assert (h, w, c) == want.shape


def _png(w, h, ctype, depth, raw_rows, level, strategy=zlib.Z_DEFAULT_STRATEGY, chunk=None):
    import struct

    def ch(t, b):
        return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b))

    co = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
    z = co.compress(raw_rows) + co.flush()
    idat = b"".join(ch(b"IDAT", z[i:i + chunk]) for i in range(0, len(z), chunk)) if chunk else ch(b"IDAT", z)
    return b"\x89PNG\r\n\x1a\n" + ch(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)) + idat + ch(b"IEND", b"")


def test_inflate_block_types_and_every_filter():
    # hand-assembled PNGs: stored / fixed-Huffman / dynamic-Huffman deflate blocks, split IDAT chunks, and every row
    # filter type applied by a reference filter implementation
    rng = np.random.default_rng(5)
    h, w = 40, 57
    img = np.kron(rng.integers(0, 256, size=(h // 4, w // 3 + 1, 3)), np.ones((4, 3, 1))).astype(np.uint8)[:, :w]
    img = (img.astype(int) + rng.integers(-3, 4, size=img.shape)).clip(0, 255).astype(np.uint8)
    bpp, stride = 3, w * 3
    rows = img.reshape(h, stride).astype(int)
    raw = bytearray()
    for y in range(h):
        f = y % 5
        cur, up = rows[y], rows[y - 1] if y else np.zeros(stride, int)
        a = np.concatenate([np.zeros(bpp, int), cur[:-bpp]])
        c = np.concatenate([np.zeros(bpp, int), up[:-bpp]])
        if f == 0:
            pred = 0
        elif f == 1:
            pred = a
        elif f == 2:
            pred = up
        elif f == 3:
            pred = (a + up) >> 1
        else:
            pa, pb, pc = abs(up - c), abs(a - c), abs(a + up - 2 * c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, up, c))
        raw.append(f)
        raw += bytes(((cur - pred) & 255).astype(np.uint8))
    for level, strategy, chunk in ((0, zlib.Z_DEFAULT_STRATEGY, None), (6, zlib.Z_FIXED, None), (9, zlib.Z_DEFAULT_STRATEGY, 100),
                                   (1, zlib.Z_HUFFMAN_ONLY, 33)):
        data = _png(w, h, 2, 8, bytes(raw), level, strategy, chunk)
        got = host_stage(data, np.uint8).reshape(h, w, 3)
        assert (got == img).all(), (level, strategy)
        assert (dec.png_decode(data) == img).all()


def test_rejected_files():
    L = _ffi.lib()
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    for bad in (b"GIF89a" + b"\0" * 32, b"\xff\xd8\xff\xd9", b"\x89PNG\r\n\x1a\n" + b"\0" * 40, b""):
        buf = np.frombuffer(bad + b"\0", np.uint8)
        assert L.ocrb_image_info(buf.ctypes.data, len(bad), C.byref(w), C.byref(h), C.byref(c)) == -1
    # truncated entropy-coded data / corrupt zlib stream: an error, not a crash
    from conftest import GOLDEN  # noqa: F401
    z = np.load(GOLDEN + "/image_files.npz")
    jpg = z["jpg_img545"].tobytes()
    need = C.c_size_t()
    for cut in (len(jpg) // 2, 700, len(jpg) - 3):
        d = np.frombuffer(jpg[:cut], np.uint8)
        L.ocrb_debug_decode_host(d.ctypes.data, cut, None, 0, C.byref(need))
        out = np.empty(need.value, np.uint8)
        rc = L.ocrb_debug_decode_host(d.ctypes.data, cut, out.ctypes.data, need.value, C.byref(need))
        assert rc in (0, -1)  # a truncated scan may still decode (zeros behind the end), but must not fault
    png = bytearray(z["synpng_rgb"].tobytes())
    png[60] ^= 0x55
    d = np.frombuffer(bytes(png), np.uint8)
    L.ocrb_debug_decode_host(d.ctypes.data, len(png), None, 0, C.byref(need))
    out = np.empty(need.value, np.uint8)
    assert L.ocrb_debug_decode_host(d.ctypes.data, len(png), out.ctypes.data, need.value, C.byref(need)) in (0, -1)
