"""CPU: the C-ABI library loads, exports every symbol include/ocrb.h declares, and refuses to
compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ffi():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "ocr_rs_b200", "libocrb.so")):
        g.build()
    from ocr_rs_b200 import _ffi
    return _ffi


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ocrb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ocrb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(ffi):
    names = _declared_symbols()
    assert len(names) >= 40
    L = ffi.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ocrb.h but not exported by libocrb.so"
        assert n in ffi.SIGNATURES, f"{n} has no ctypes signature in _ffi.py"
    assert sorted(ffi.SIGNATURES) == names
    assert L.ocrb_version() == 100


def test_host_only_helpers(ffi):
    L = ffi.lib()
    rw, rh = C.c_int(), C.c_int()
    # image_ops.rs:805-1008 fixtures: 300x200 -> 800x533, 184x274 -> 537x800 (integer floor)
    assert L.ocrb_resize_dims(300, 200, 800, 800, C.byref(rw), C.byref(rh)) == 0 and (rw.value, rh.value) == (800, 533)
    assert L.ocrb_resize_dims(184, 274, 800, 800, C.byref(rw), C.byref(rh)) == 0 and (rw.value, rh.value) == (537, 800)
    assert L.ocrb_resize_dims(1, 5000, 800, 800, C.byref(rw), C.byref(rh)) == 0 and (rw.value, rh.value) == (1, 800)
    from ocr_rs_b200 import utils
    assert "".join(utils.class_to_char(i) for i in range(62)) == utils.VALUES
    assert utils.class_to_char(62) == "?"
    assert utils.parse_dimensions("800x600") == (800, 600)
    for bad in ("800", "800x600x3", "ax600", "800X600"):  # utils.rs:72-79: two 'x'-separated u32 values or an error
        with pytest.raises(ValueError):
            utils.parse_dimensions(bad)
    # utils::topk (utils.rs:28-43): largest first, class -> char through POS_TO_CHAR, the three accepted shapes
    import numpy as np
    p = np.zeros(62)
    p[[3, 30, 61]] = (0.2, 0.5, 0.3)
    assert utils.topk(p, 1) == [("e", 0.5)]
    assert utils.topk(p.reshape(1, 62), 3) == [("e", 0.5), ("9", 0.3), ("D", 0.2)]
    assert utils.topk(p.reshape(1, 1, 62), 2) == [("e", 0.5), ("9", 0.3)]
    with pytest.raises(ValueError):
        utils.topk(np.zeros((2, 62)), 1)
    assert utils.VALUES_MAP["A"] == 0 and utils.VALUES_MAP["9"] == 61 and utils.POS_TO_CHAR[26] == "a"
    assert utils.parse_number("12", "width") == 12
    with pytest.raises(ValueError):
        utils.parse_number("twelve", "width")
    p = ffi.PostprocParams()
    L.ocrb_postproc_default_params(C.byref(p))
    assert (p.thresh, p.box_thresh, p.min_size, p.unclip_factor) == (0.6, 0.7, 5.0, 2.0)  # metrics.rs:38,64,66,103


def test_no_cpu_fallback(ffi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(ffi.OcrbError) as e:
        ffi.Context(0)
    assert e.value.code == -2  # OCRB_ERR_CUDA


def test_rust_sys_crate_is_in_sync_with_the_header():
    """rust/ocrb-sys/src/lib.rs is generated from include/ocrb.h (tools/gen_rust_bindings.py): every declared symbol has an
    `extern "C"` declaration and the committed file is what the generator emits today."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.call([sys.executable, os.path.join(root, "tools", "gen_rust_bindings.py"), "--check"]) == 0
    rs = open(os.path.join(root, "rust", "ocrb-sys", "src", "lib.rs")).read()
    assert sorted(re.findall(r"pub fn (ocrb_[a-z_0-9]+)\(", rs)) == _declared_symbols()
