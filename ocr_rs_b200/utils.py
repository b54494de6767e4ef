"""utils.rs mirror: class table and top-k decoding."""
from . import _ffi

VALUES = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789"  # utils.rs:7
VALUES_COUNT = len(VALUES)


def class_to_char(cls: int) -> str:
    """utils::POS_TO_CHAR (utils.rs:11-18) through the C ABI."""
    return _ffi.lib().ocrb_class_to_char(int(cls)).decode()


def parse_dimensions(s: str):
    """utils::parse_dimensions ("800x800", utils.rs:72-79)."""
    w, h = s.lower().split("x")
    return int(w), int(h)
