"""ORACLE (test infrastructure only): the file decode in front of image_ops::preprocess_image / load_image_as_tensor
(`image::open(file)?.into_rgba()` / `.into_luma()`, image_ops.rs:193 and :78).

JPEG: ctypes front end of oracle/decode_oracle.c (jpeg-decoder 0.1.20 restated; pinned bit for bit on the reference's
four preprocessed_img*.png fixtures, see that file's header).  PNG (`png` 0.16.7): lossless, so any conforming
decoder is the oracle — zlib inflate + the five PNG filters, 8-bit and sub-byte grey / palette / RGB / RGBA,
non-interlaced.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdecode_oracle.so")
_lib = None

JPEG_COLOR_F32 = 0    # ycbcr_to_rgb in f32 with +0.5 truncation: the form that reproduces the reference's fixtures
JPEG_COLOR_FIXED = 1  # 20-bit fixed point (later jpeg-decoder releases); differs on the img55 fixture


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "decode_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "libdecode_oracle.so"])
        _lib = C.CDLL(_SO)
    return _lib


def jpeg_decode(data: bytes, variant: int = JPEG_COLOR_F32) -> np.ndarray:
    """-> uint8 [h, w, 1 | 3] (L8 or RGB8, jpeg-decoder's PixelFormat)."""
    data = bytes(data)
    px = C.POINTER(C.c_uint8)()
    w, h, nc = C.c_int(), C.c_int(), C.c_int()
    rc = lib().orc_jpeg_decode(data, C.c_size_t(len(data)), int(variant), C.byref(px), C.byref(w), C.byref(h), C.byref(nc))
    if rc != 0:
        raise ValueError(f"jpeg decode failed ({'unsupported' if rc == -2 else 'malformed'})")
    a = np.ctypeslib.as_array(px, shape=(h.value, w.value, nc.value)).copy()
    lib().orc_decode_free(px)
    return a


def jpeg_coefficients(data: bytes) -> np.ndarray:
    """The entropy-decoded DCT coefficients (int16, components concatenated, natural order, MCU-padded block grid)."""
    data = bytes(data)
    px, co = C.POINTER(C.c_uint8)(), C.POINTER(C.c_int16)()
    w, h, nc, cnt = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
    rc = lib().orc_jpeg_decode_ex(data, C.c_size_t(len(data)), 0, C.byref(px), C.byref(w), C.byref(h), C.byref(nc), C.byref(co), C.byref(cnt))
    if rc != 0:
        raise ValueError("jpeg decode failed")
    a = np.ctypeslib.as_array(co, shape=(cnt.value,)).copy()
    lib().orc_decode_free(px)
    lib().orc_decode_free(co)
    return a


def png_decode(data: bytes) -> np.ndarray:
    """-> uint8 [h, w, c], c = 1 (L), 2 (LA), 3 (RGB) or 4 (RGBA) after the expansions `png` 0.16 applies for the
    image crate (palette -> RGB, tRNS -> alpha, sub-byte grey -> 8 bits)."""
    data = bytes(data)
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG file")
    pos, idat, plte, trns, hdr = 8, [], None, None, None
    while pos + 8 <= len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"PLTE":
            plte = np.frombuffer(body, np.uint8).reshape(-1, 3)
        elif typ == b"tRNS":
            trns = np.frombuffer(body, np.uint8)
        elif typ == b"IDAT":
            idat.append(body)
        elif typ == b"IEND":
            break
    w, h, depth, ctype, _, _, interlace = hdr
    if interlace or depth == 16:
        raise ValueError("unsupported PNG (interlaced or 16-bit)")
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    bpp = max(1, ch * depth // 8)
    stride = (w * ch * depth + 7) // 8
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), np.uint8)
    rows = np.zeros((h + 1, stride), np.int32)
    for y in range(h):
        f = int(raw[y * (stride + 1)])
        line = raw[y * (stride + 1) + 1:(y + 1) * (stride + 1)].astype(np.int32)
        up = rows[y]
        cur = rows[y + 1]
        if f == 0:
            cur[:] = line
        elif f == 2:
            cur[:] = (line + up) & 255
        else:
            for x in range(stride):
                a = cur[x - bpp] if x >= bpp else 0
                b = up[x]
                c = up[x - bpp] if x >= bpp else 0
                if f == 1:
                    p = a
                elif f == 3:
                    p = (a + b) >> 1
                else:
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    p = a if pa <= pb and pa <= pc else (b if pb <= pc else c)
                cur[x] = (line[x] + p) & 255
    pix = rows[1:].astype(np.uint8)
    if depth < 8:  # unpack sub-byte samples, most significant first
        per = 8 // depth
        shifts = np.arange(per - 1, -1, -1) * depth
        pix = ((pix[:, :, None] >> shifts) & ((1 << depth) - 1)).reshape(h, -1)[:, :w * ch]
        if ctype == 0:
            pix = (pix * (255 // ((1 << depth) - 1))).astype(np.uint8)
    pix = pix.reshape(h, w, ch)
    if ctype == 3:
        idx = pix[..., 0]
        out = plte[idx]
        if trns is not None:
            alpha = np.full(256, 255, np.uint8)
            alpha[:len(trns)] = trns
            out = np.concatenate([out, alpha[idx][..., None]], 2)
        return np.ascontiguousarray(out)
    if trns is not None and ctype == 0:
        key = struct.unpack(">H", bytes(trns[:2]))[0]
        key8 = key * (255 // ((1 << depth) - 1)) if depth < 8 else key
        alpha = np.where(pix[..., 0] == key8, 0, 255).astype(np.uint8)
        return np.concatenate([pix, alpha[..., None]], 2)
    if trns is not None and ctype == 2:
        key = np.array(struct.unpack(">HHH", bytes(trns[:6])), np.uint8)
        alpha = np.where((pix == key).all(-1), 0, 255).astype(np.uint8)
        return np.concatenate([pix, alpha[..., None]], 2)
    return np.ascontiguousarray(pix)


def to_rgba(px: np.ndarray) -> np.ndarray:
    """DynamicImage::into_rgba (image 0.23.11): L -> (l, l, l, 255), LA -> (l, l, l, a), RGB -> (r, g, b, 255)."""
    h, w, c = px.shape
    out = np.empty((h, w, 4), np.uint8)
    if c == 1:
        out[..., :3] = px
        out[..., 3] = 255
    elif c == 2:
        out[..., :3] = px[..., :1]
        out[..., 3] = px[..., 1]
    elif c == 3:
        out[..., :3] = px
        out[..., 3] = 255
    else:
        out[:] = px
    return out


def to_luma(px: np.ndarray) -> np.ndarray:
    """DynamicImage::into_luma (image 0.23.11 color.rs): Rec.709 weights in f32, truncating cast
    (the same conversion preprocess_image applies after the resize)."""
    h, w, c = px.shape
    if c <= 2:
        return np.ascontiguousarray(px[..., 0])
    r, g, b = (px[..., i].astype(np.float32) for i in range(3))
    l = np.float32(0.2126) * r + np.float32(0.7152) * g + np.float32(0.0722) * b
    return l.astype(np.uint8)


def open_image(data: bytes) -> np.ndarray:
    """image::open by content: JPEG or PNG -> decoded pixels [h, w, c]."""
    data = bytes(data)
    if data[:2] == b"\xff\xd8":
        return jpeg_decode(data)
    if data[:8] == b"\x89PNG\r\n\x1a\n":
        return png_decode(data)
    raise ValueError("unsupported image format")
