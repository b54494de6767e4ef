"""ocr_rs_b200 — B200-native replacement for the text-detection / glyph-recognition
inference hot path of lazareviczoran/ocr-rs.

The product is libocrb.so (hand-written sm_100a CUDA behind the C ABI in include/ocrb.h).
This package is the host-side mirror of the reference's module entry points
(image_ops / text_detection::{model,metrics} / polygon / char_recognition) on top of that
ABI, used by the tests and bench.py; see INTEGRATION.md for the Rust binding.
There is no CPU fallback anywhere in this package.
"""
from . import _ffi  # noqa: F401
from ._ffi import Context, OcrbError, default_context  # noqa: F401

__all__ = ["Context", "OcrbError", "default_context", "image_ops", "polygon", "text_detection",
           "char_recognition", "utils", "synth"]
