"""The C++ host façade (include/qrkit_b200/QRKit.hpp): compiles against the C ABI with g++, and — on the GPU box —
passes the reference's own block-diagonal test properties.  On a CPU-only box it must report "no device" (exit 77),
never compute."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from qrkit_b200.build import build_library
    lib = build_library()
    exe = str(tmp_path / "test_facade")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_facade.cpp"),
                    "-o", exe, lib, f"-Wl,-rpath,{os.path.dirname(lib)}"], check=True)
    return exe


def test_facade_compiles_and_refuses_without_gpu(tmp_path):
    from qrkit_b200 import capi
    exe = _build(tmp_path)
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present; the gpu test runs the program")
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 77, res.stdout + res.stderr
    assert "no CPU fallback" in res.stdout


@pytest.mark.gpu
def test_facade_reference_properties_on_gpu(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "All passed." in res.stdout and "FAILED" not in res.stdout
