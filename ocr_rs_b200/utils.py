"""utils.rs mirror: class table, top-k decoding, argument parsing (utils.rs:7-79)."""
import numpy as np

from . import _ffi

VALUES = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789"  # utils.rs:7
VALUES_COUNT = len(VALUES)
POS_TO_CHAR = {i: ch for i, ch in enumerate(VALUES)}  # utils.rs:12-18
VALUES_MAP = {ch: i for i, ch in enumerate(VALUES)}   # utils.rs:19-25


def class_to_char(cls: int) -> str:
    """utils::POS_TO_CHAR (utils.rs:11-18) through the C ABI."""
    return _ffi.lib().ocrb_class_to_char(int(cls)).decode()


def topk(tensor, k: int):
    """utils::topk (utils.rs:28-43): the k largest entries of a [62] / [1, 62] / [1, 1, 62] vector of class scores
    (run_prediction hands it softmax(-1, Double) of the logits, char_recognition/mod.rs:53-56) as [(char, value)],
    largest first.  Any other shape is the reference's panic -> ValueError.  Ties keep the lower class first (what
    libtorch's sorted CPU topk does for equal values)."""
    t = np.asarray(tensor, np.float64)
    if t.shape not in ((VALUES_COUNT,), (1, VALUES_COUNT), (1, 1, VALUES_COUNT)):
        raise ValueError(f"unexpected tensor shape {list(t.shape)}")
    t = t.reshape(VALUES_COUNT)
    if not 0 <= k <= VALUES_COUNT:
        raise ValueError(f"k = {k} out of range for {VALUES_COUNT} classes")
    order = np.argsort(-t, kind="stable")[:k]
    return [(POS_TO_CHAR[int(i)], float(t[i])) for i in order]


def parse_number(num_str: str, field: str, kind=int):
    """utils::parse_number (utils.rs:65-70)."""
    try:
        return kind(num_str)
    except ValueError:
        raise ValueError(f"Could not parse {field} value: {num_str}") from None


def parse_dimensions(s: str):
    """utils::parse_dimensions ("800x800", utils.rs:72-79): exactly two 'x'-separated unsigned integers."""
    values = s.split("x")
    if values and values[-1] == "":  # split_terminator: a trailing separator yields no empty last piece
        values = values[:-1]
    if len(values) != 2:
        raise ValueError(f"Could not parse dimensions value: {s}")
    w, h = (int(v) for v in values)
    if w < 0 or h < 0 or w > 0xFFFFFFFF or h > 0xFFFFFFFF:
        raise ValueError(f"Could not parse dimensions value: {s}")
    return w, h
