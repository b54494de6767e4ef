"""The C++ host mirror of ocr-rs's modules (include/ocrb.hpp) compiles against the C ABI and runs from compiled code:
host-only behaviour on CPU, the reference's known-answer tests through the mirror on a GPU box."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import __graft_entry__ as g
    lib_dir = os.path.join(ROOT, "ocr_rs_b200")
    if not os.path.exists(os.path.join(lib_dir, "libocrb.so")):
        g.build()
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = str(tmp_path / "host_mirror_test")
    subprocess.run([gxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
                    "-L", lib_dir, "-l:libocrb.so", "-Wl,-rpath," + lib_dir, "-o", exe], check=True)
    return exe


def test_cpp_mirror_compiles_and_runs_host_only(tmp_path):
    import torch
    exe = _build(tmp_path)
    out = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "varstore_libtorch.ot")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    if not torch.cuda.is_available():
        assert "no CUDA device" in out.stdout


@pytest.mark.gpu
def test_cpp_mirror_on_device(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "varstore_libtorch.ot")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "device checks ok" in out.stdout
