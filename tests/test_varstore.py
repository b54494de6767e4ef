"""Model files: the native reader of tch's VarStore archives (csrc/varstore.cu) against a file
libtorch itself wrote (tests/golden/varstore_libtorch.ot, generator committed beside it) and
against libtorch's own reader (torch.jit.load); the Python writer round trip."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "varstore_libtorch.ot")


def test_native_reader_on_libtorch_written_archive():
    torch = pytest.importorskip("torch")
    from ocr_rs_b200 import varstore
    got = varstore.load_varstore(FIXTURE)
    ref = dict(torch.jit.load(FIXTURE).named_parameters())
    assert list(got) == list(ref) and len(got) == 18  # same names, VarStore order kept
    for k, v in ref.items():
        a = v.detach().to(torch.float32).numpy()
        assert got[k].shape == a.shape and (got[k] == a).all(), k
    assert got["transposed_view"].tolist() == np.arange(12, dtype=np.float32).reshape(3, 4).T.tolist()
    assert got["offset_view"].ravel().tolist() == list(range(5, 15))
    assert got["scalar"].shape == () and float(got["scalar"]) == 3.25


def test_python_writer_round_trip_and_libtorch_accepts_it(tmp_path):
    torch = pytest.importorskip("torch")
    from ocr_rs_b200 import synth, varstore
    for weights in (synth.make_rec_weights(1), synth.make_detector_weights(0, "hard_bn")):  # 8 and 121 tensors
        path = str(tmp_path / "model.ot")
        varstore.save_varstore(weights, path)
        back = varstore.load_varstore(path)
        assert list(back) == list(weights)
        for k, v in weights.items():
            assert back[k].shape == v.shape and (back[k] == v).all(), k
        theirs = dict(torch.jit.load(path).named_parameters())  # libtorch's own reader takes the file
        assert set(theirs) == set(weights)
        for k in list(weights)[::7]:
            assert (theirs[k].detach().numpy() == weights[k]).all()


def test_reader_errors(tmp_path):
    from ocr_rs_b200 import OcrbError, varstore
    with pytest.raises(OcrbError):
        varstore.load_varstore(str(tmp_path / "missing.ot"))
    bad = tmp_path / "garbage.ot"
    bad.write_bytes(b"this is not a zip archive" * 10)
    with pytest.raises(OcrbError):
        varstore.load_varstore(str(bad))
    trunc = tmp_path / "truncated.ot"
    trunc.write_bytes(open(FIXTURE, "rb").read()[:3000])
    with pytest.raises(OcrbError):
        varstore.load_varstore(str(trunc))


@pytest.mark.gpu
def test_create_from_file_matches_create_from_arrays(tmp_path):
    """vs.load(file) path: ocrb_det_create_from_file / ocrb_rec_create_from_file give the same
    nets as the array entry points (bit-identical outputs)."""
    import ctypes as C
    from ocr_rs_b200 import _ffi, synth, varstore
    from ocr_rs_b200.char_recognition.model import Net
    from ocr_rs_b200.text_detection.model import FuncT, resnet18
    ctx = _ffi.default_context()
    wd, wr = synth.make_detector_weights(5, "hard_bn"), synth.make_rec_weights(6)
    pd, pr = str(tmp_path / "text_detection.model"), str(tmp_path / "char_rec_conv_net.model")
    varstore.save_varstore(wd, pd)
    aliased = {a: wr[n] for (n, _), a in zip(synth.REC_CANONICAL, synth.REC_VARSTORE_ALIASES)}  # what tch really writes
    varstore.save_varstore(aliased, pr)
    x = synth.make_noise_images(2, 96, 128, seed=1).reshape(2, 1, 96, 128)
    g = synth.make_glyphs(33, 2, "strokes")
    for mode, code in (("bf16", _ffi.MODE_BF16), ("fp32", _ffi.MODE_FP32)):
        a = resnet18(wd, mode).forward_t(x)
        net = FuncT.__new__(FuncT)
        net.ctx, net.mode, net._h = ctx, code, _ffi.c_p()
        _ffi.check(_ffi.lib().ocrb_det_create_from_file(ctx.handle, pd.encode(), code, C.byref(net._h)))
        assert (net.forward_t(x) == a).all()
    rec = Net.__new__(Net)
    rec.ctx, rec._h = ctx, _ffi.c_p()
    _ffi.check(_ffi.lib().ocrb_rec_create_from_file(ctx.handle, pr.encode(), C.byref(rec._h)))
    assert (rec.predict(g)[0] == Net(wr).predict(g)[0]).all()


def test_corrupt_archives_are_rejected_not_fatal(tmp_path):
    """ocrb.h: nothing throws across the ABI.  Truncated files, absurd sizes in the pickle and wrapped ZIP
    offsets must come back as OCRB_ERR_INVALID (host-only: no device needed)."""
    import ctypes as C
    import os

    from ocr_rs_b200 import _ffi
    src = open(os.path.join(os.path.dirname(__file__), "golden", "varstore_libtorch.ot"), "rb").read()
    L = _ffi.lib()

    def opens(blob):
        p = tmp_path / "m.ot"
        p.write_bytes(blob)
        h = _ffi.c_p()
        rc = L.ocrb_varstore_open(str(p).encode(), C.byref(h))
        if rc == 0:
            L.ocrb_varstore_close(h)
        return rc

    assert opens(src) == 0
    assert opens(src[:10]) == -1 and opens(b"") == -1 and opens(src[: len(src) // 2]) == -1
    rng = np.random.default_rng(0)
    for _ in range(300):  # random byte corruption anywhere (directory, pickle, sizes): never a crash
        b = bytearray(src)
        for _ in range(int(rng.integers(1, 8))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        assert opens(bytes(b)) in (0, -1)
    # a huge dimension written into the pickle's size tuple (BININT 'J' + 4 bytes little endian)
    b = bytearray(src)
    pk = src.find(b"data.pkl")
    for i in range(pk, len(b) - 5):
        if b[i] == ord("J"):
            b[i + 1:i + 5] = (0x7fffffff).to_bytes(4, "little")
    assert opens(bytes(b)) in (0, -1)
