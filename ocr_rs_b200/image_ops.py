"""image_ops.rs mirror (inference subset): same names, argument meaning and error behaviour.

preprocess_image            image_ops.rs:188-220
convert_image_to_tensor     image_ops.rs:350-364
convert_tensor_to_image     image_ops.rs:367-381
load_image_as_tensor        image_ops.rs:73-85
File decoding (image::open): preprocess_image / load_image_as_tensor take a file path or the
encoded bytes like the reference (JPEG / PNG, decoded by libocrb: entropy decode on host threads, IDCT + upsampling +
colour conversion on the device) — or, as before, already decoded pixel arrays.
"""
import ctypes as C
import os

import numpy as np

from . import _ffi


def _ctx(ctx):
    return ctx if ctx is not None else _ffi.default_context()


def _file_bytes(src):
    """path / bytes -> bytes; raises like image::open for a missing file (image_ops.rs:193)."""
    if isinstance(src, (bytes, bytearray, memoryview)):
        return bytes(src)
    with open(os.fspath(src), "rb") as f:
        return f.read()


def _blobs(files):
    blobs = [_file_bytes(f) for f in files]
    n = len(blobs)
    keep = [np.frombuffer(b, np.uint8) for b in blobs]
    ptrs = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
    sizes = (C.c_size_t * n)(*[len(b) for b in blobs])
    return keep, ptrs, sizes, n


def image_info(file):
    """(width, height, channels) of an encoded JPEG / PNG file (path or bytes) — host only."""
    b = _file_bytes(file)
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    buf = np.frombuffer(b, np.uint8)
    _ffi.check(_ffi.lib().ocrb_image_info(buf.ctypes.data, len(b), C.byref(w), C.byref(h), C.byref(c)))
    return w.value, h.value, c.value


def decode_images(files, fmt="rgba", ctx=None):
    """image::open(file)?.into_rgba() / .into_luma() for a list of files (paths or bytes)
    -> list of uint8 arrays [h, w, 4] (fmt="rgba") or [h, w] (fmt="luma")."""
    ctx = _ctx(ctx)
    keep, ptrs, sizes, n = _blobs(files)
    bpp = 4 if fmt == "rgba" else 1
    dims = [image_info(k.tobytes())[:2] for k in keep]
    offs = np.zeros(n, np.int64)
    offs[1:] = np.cumsum([w * h * bpp for w, h in dims[:-1]])
    out = np.empty(int(offs[-1]) + dims[-1][0] * dims[-1][1] * bpp, np.uint8)
    _ffi.check(_ffi.lib().ocrb_decode_images(ctx.handle, ptrs, sizes, n, _ffi.PIXELS_RGBA if fmt == "rgba" else _ffi.PIXELS_LUMA,
                                             _ffi.ptr(offs), _ffi.ptr(out)))
    res = []
    for (w, h), o in zip(dims, offs):
        a = out[o:o + w * h * bpp]
        res.append(a.reshape(h, w, 4) if fmt == "rgba" else a.reshape(h, w))
    return res


def preprocess_files(files, target_dim, ctx=None, out=None):
    """preprocess_image(file, target_dim) for a batch of encoded files (paths or bytes) in one call
    (ocrb_preprocess_files: the decoded pixels stay in HBM) -> (uint8 [n, height, width], adjust float64 [n, 2])."""
    ctx = _ctx(ctx)
    W, H = int(target_dim[0]), int(target_dim[1])
    keep, ptrs, sizes, n = _blobs(files)
    if out is None:
        out = np.empty((n, H, W), np.uint8)
    adj = np.empty((n, 2), np.float64)
    _ffi.check(_ffi.lib().ocrb_preprocess_files(ctx.handle, ptrs, sizes, n, W, H, _ffi.ptr(out), _ffi.ptr(adj)))
    return out, adj


def preprocess_image(image, target_dim, ctx=None):
    """image: a file path / encoded bytes (the reference's signature, image_ops.rs:188) or decoded
    uint8 [h, w, 4] pixels (DynamicImage::into_rgba); target_dim = (width, height)
    -> (GrayImage uint8 [height, width], adjust_x, adjust_y)."""
    ctx = _ctx(ctx)
    if not isinstance(image, np.ndarray):
        out, adj = preprocess_files([image], target_dim, ctx)
        return out[0], float(adj[0, 0]), float(adj[0, 1])
    rgba = image
    rgba = np.ascontiguousarray(rgba, np.uint8)
    if rgba.ndim != 3 or rgba.shape[2] != 4:
        raise ValueError("expected an RGBA8 image [h, w, 4]")
    W, H = int(target_dim[0]), int(target_dim[1])
    sh, sw = rgba.shape[:2]
    out = np.empty((H, W), np.uint8)
    ax, ay = C.c_double(), C.c_double()
    _ffi.check(_ffi.lib().ocrb_preprocess_rgba(ctx.handle, _ffi.ptr(rgba), sw, sh, W, H, _ffi.ptr(out),
                                               C.byref(ax), C.byref(ay)))
    return out, ax.value, ay.value


def preprocess_images(rgbas, target_dim, ctx=None, out=None):
    """preprocess_image for a batch in one fused launch (ocrb_preprocess_rgba_batch).
    rgbas: list of uint8 [h_i, w_i, 4] arrays, or a tuple (packed uint8 buffer (numpy / torch, host or cuda), offsets int64 [n],
    widths int32 [n], heights int32 [n]).  -> (uint8 [n, height, width], adjust float64 [n, 2])"""
    ctx = _ctx(ctx)
    W, H = int(target_dim[0]), int(target_dim[1])
    if isinstance(rgbas, tuple):
        buf, offs, ws, hs = rgbas
        offs = np.ascontiguousarray(offs, np.int64)
        ws, hs = np.ascontiguousarray(ws, np.int32), np.ascontiguousarray(hs, np.int32)
    else:
        arrs = [np.ascontiguousarray(a, np.uint8) for a in rgbas]
        for a in arrs:
            if a.ndim != 3 or a.shape[2] != 4:
                raise ValueError("expected RGBA8 images [h, w, 4]")
        offs = np.zeros(len(arrs), np.int64)
        offs[1:] = np.cumsum([a.size for a in arrs[:-1]])
        buf = np.concatenate([a.reshape(-1) for a in arrs])
        ws = np.array([a.shape[1] for a in arrs], np.int32)
        hs = np.array([a.shape[0] for a in arrs], np.int32)
    n = len(offs)
    if out is None:
        out = np.empty((n, H, W), np.uint8)
    adj = np.empty((n, 2), np.float64)
    _ffi.check(_ffi.lib().ocrb_preprocess_rgba_batch(ctx.handle, _ffi.ptr(buf), _ffi.ptr(offs), _ffi.ptr(ws), _ffi.ptr(hs), n, W, H,
                                                     _ffi.ptr(out), _ffi.ptr(adj)))
    return out, adj


def convert_image_to_tensor(image, ctx=None, out=None):
    """GrayImage uint8 [H, W] -> float32 tensor [H, W] (the reference builds f64 then
    .to_kind(Float); values are 0..255, no scaling)."""
    ctx = _ctx(ctx)
    image = np.ascontiguousarray(image, np.uint8)
    if out is None:
        out = np.empty(image.shape, np.float32)
    _ffi.check(_ffi.lib().ocrb_convert_image_to_tensor(ctx.handle, _ffi.ptr(image), image.size, _ffi.ptr(out)))
    return out


def convert_tensor_to_image(tensor, scale=1.0, ctx=None):
    """float32 [H, W] -> GrayImage uint8 by truncation; errors on > 2 dims like the reference."""
    ctx = _ctx(ctx)
    tensor = np.ascontiguousarray(tensor, np.float32)
    if tensor.ndim > 2:
        raise ValueError("tensor must be in 2 dimensions")  # image_ops.rs:369-371
    out = np.empty(tensor.shape, np.uint8)
    _ffi.check(_ffi.lib().ocrb_convert_tensor_to_image(ctx.handle, _ffi.ptr(tensor), tensor.size, float(scale), _ffi.ptr(out)))
    return out


def load_image_as_tensor(luma, ctx=None):
    """file path / encoded bytes (image_ops.rs:73: open(file)?.into_luma()) or luma uint8 [h, w]
    -> float32 [1, w*h] = pixel / 255 (image_ops.rs:79-83)."""
    ctx = _ctx(ctx)
    if not isinstance(luma, np.ndarray):
        if isinstance(luma, (str, os.PathLike)) and not os.path.exists(luma):
            raise FileNotFoundError(f"File {luma} doesn't exist")  # image_ops.rs:75-77
        luma = decode_images([luma], "luma", ctx)[0]
    luma = np.ascontiguousarray(luma, np.uint8)
    out = np.empty((1, luma.size), np.float32)
    _ffi.check(_ffi.lib().ocrb_load_image_as_tensor(ctx.handle, _ffi.ptr(luma), luma.size, _ffi.ptr(out)))
    return out
