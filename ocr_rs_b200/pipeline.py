"""Whole-path entry points above the C ABI (include/ocrb.h "pipeline"): the device part of run_text_detection for a
batch (text_detection/mod.rs:46-67, :188-204) with recognition in the same call, and the polygon -> glyph crop glue
the reference leaves open (README.md:20-26)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi


def detect_and_recognize(det, rec, images, adjust, glyphs=None, params=None):
    """images u8 [B,H,W] (numpy or torch, host or cuda), adjust f64 [B,2], glyphs u8 [n,784] caller-provided crops
    -> (_ffi.Polygons, argmax int32 [n] or None)"""
    B, H, W = images.shape
    adjust = np.ascontiguousarray(adjust, np.float64)
    n = 0 if glyphs is None else len(glyphs)
    am = np.empty(n, np.int32) if n else None
    h = _ffi.c_p()
    _ffi.check(_ffi.lib().ocrb_detect_and_recognize(det._h, rec._h if rec is not None else None, _ffi.ptr(images), _ffi.ptr(adjust),
                                                    B, H, W, params, _ffi.ptr(glyphs), n, _ffi.ptr(am), C.byref(h)))
    return _ffi.Polygons(h), am


def detect_and_read(det, rec, images, adjust, glyphs_per_polygon=4, params=None):
    """detector -> post-processing -> crop glue -> recognition in one call: the polygons plus
    .glyph_classes int32 [n_polygons, glyphs_per_polygon] (utils.class_to_char maps a class to its character)."""
    B, H, W = images.shape
    adjust = np.ascontiguousarray(adjust, np.float64)
    h = _ffi.c_p()
    _ffi.check(_ffi.lib().ocrb_detect_and_read(det._h, rec._h, _ffi.ptr(images), _ffi.ptr(adjust), B, H, W, params,
                                               int(glyphs_per_polygon), C.byref(h)))
    return _ffi.Polygons(h)


def crop_glyphs(image, boxes, glyphs_per_box, ctx=None):
    """crop spec v1 (csrc/crop.cu): image u8 [H,W], boxes int32 [n,4,2] (TL,TR,BR,BL) -> u8 [n*glyphs_per_box, 784]"""
    ctx = ctx if ctx is not None else _ffi.default_context()
    image = np.ascontiguousarray(image, np.uint8) if not hasattr(image, "data_ptr") else image
    b = np.ascontiguousarray(np.asarray(boxes, np.int32).reshape(-1, 4, 2))
    out = np.empty((len(b) * glyphs_per_box, 784), np.uint8)
    _ffi.check(_ffi.lib().ocrb_crop_glyphs(ctx.handle, _ffi.ptr(image), int(image.shape[0]), int(image.shape[1]), _ffi.ptr(b), len(b),
                                           int(glyphs_per_box), _ffi.ptr(out)))
    return out
