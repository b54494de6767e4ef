"""VarStore model files (utils.rs:55-63 `save_vs`, text_detection/mod.rs:40-44 `vs.load`).

load_varstore   native reader (libocrb, csrc/varstore.cu): name -> float32 array
save_varstore   writes the same libtorch archive layout tch 0.3.0 produces through
                torch::serialize::OutputArchive (ZIP of STORED entries: data.pkl + data/<n> +
                code/__torch__.py + constants.pkl + version), so a weight dict made here can be
                loaded by the reference (`--model-file`) and by torch.jit.load.  Pure Python.
"""
import ctypes as C
import struct
import zipfile

import numpy as np

from . import _ffi


def load_varstore(path):
    L = _ffi.lib()
    h = _ffi.c_p()
    _ffi.check(L.ocrb_varstore_open(str(path).encode(), C.byref(h)))
    try:
        out = {}
        for i in range(L.ocrb_varstore_count(h)):
            data, numel, shape, ndim = C.POINTER(C.c_float)(), _ffi.i64(), C.POINTER(_ffi.i64)(), C.c_int()
            _ffi.check(L.ocrb_varstore_tensor(h, i, C.byref(data), C.byref(numel), C.byref(shape), C.byref(ndim)))
            shp = tuple(int(shape[k]) for k in range(ndim.value))
            arr = np.ctypeslib.as_array(data, shape=(numel.value,)).copy() if numel.value else np.zeros(0, np.float32)
            out[L.ocrb_varstore_name(h, i).decode()] = arr.reshape(shp)
        return out
    finally:
        L.ocrb_varstore_close(h)


def _pickle_module(weights):
    """Protocol-2 pickle of `__torch__.Module` whose state is {name: _rebuild_tensor_v2(...)} —
    opcode for opcode what libtorch's pickler emits for an OutputArchive."""
    out = bytearray(b"\x80\x02c__torch__\nModule\nq\x00)\x81}(")
    memo = 1

    def put():
        nonlocal memo
        b = b"q" + bytes([memo]) if memo < 256 else b"r" + struct.pack("<I", memo)
        memo += 1
        return b

    def uni(s):
        e = s.encode()
        return b"X" + struct.pack("<I", len(e)) + e

    def integer(v):
        if 0 <= v < 256:
            return b"K" + bytes([v])
        if 0 <= v < 65536:
            return b"M" + struct.pack("<H", v)
        return b"J" + struct.pack("<i", v)

    first = True
    ids = {}
    for k, (name, arr) in enumerate(weights.items()):
        out += uni(name) + put()
        if first:
            out += b"ctorch._utils\n_rebuild_tensor_v2\n"
            ids["rebuild"] = memo
            out += put()
            out += b"((" + uni("storage")
            ids["storage"] = memo
            out += put() + b"ctorch\nFloatStorage\n"
            ids["float"] = memo
            out += put()
        else:
            out += b"h" + bytes([ids["rebuild"]]) + b"((h" + bytes([ids["storage"]]) + b"h" + bytes([ids["float"]])
        out += uni(str(k)) + put()
        if first:
            out += uni("cpu")
            ids["cpu"] = memo
            out += put()
        else:
            out += b"h" + bytes([ids["cpu"]])
        out += integer(arr.size) + b"tQ" + put() + b"K\x00("
        for d in arr.shape:
            out += integer(int(d))
        out += b"t("
        stride = [int(np.prod(arr.shape[i + 1:])) for i in range(arr.ndim)]
        for d in stride:
            out += integer(d)
        out += b"t\x89"
        if first:
            out += b"ccollections\nOrderedDict\n"
            ids["od"] = memo
            out += put()
        else:
            out += b"h" + bytes([ids["od"]])
        out += b")RtR"
        first = False
    out += b"ub" + put() + b"."
    return bytes(out)


def save_varstore(weights, path, root="archive"):
    """weights: ordered dict name -> float32 array (any shape)."""
    weights = {k: np.ascontiguousarray(v, np.float32) for k, v in weights.items()}
    code = "class Module(Module):\n  __parameters__ = [" + "".join(f'"{n}", ' for n in weights) + "]\n  __buffers__ = []\n  __annotations__ = []\n"
    code += "".join(f'  __annotations__["{n}"] = Tensor\n' for n in weights)
    with zipfile.ZipFile(path, "w", zipfile.ZIP_STORED, allowZip64=True) as z:
        for k, arr in enumerate(weights.values()):
            z.writestr(f"{root}/data/{k}", arr.tobytes())
        z.writestr(f"{root}/data.pkl", _pickle_module(weights))
        z.writestr(f"{root}/code/__torch__.py", code)
        z.writestr(f"{root}/constants.pkl", b"\x80\x02).")
        z.writestr(f"{root}/version", b"3\n")
