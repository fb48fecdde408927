"""include/qrkit_b200/EigenAdapter.hpp — the reference-side binding (Eigen::SparseSolverBase CRTP classes for the three solvers
plus the HasRowsPermutation trait).  Eigen is not in this image, so the header is compiled against a minimal mock of the Eigen and
QRKit names it touches (tests/cpp/mock_eigen, tests/cpp/mock_qrkit: test infrastructure only): every member is instantiated, the
program links against the C ABI, and without a device compute() must report InvalidInput (exit 77)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from qrkit_b200.build import build_library
    lib = build_library()
    exe = str(tmp_path / "test_eigen_adapter")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unused-local-typedefs",
                    "-I", os.path.join(ROOT, "tests", "cpp", "mock_eigen"), "-I", os.path.join(ROOT, "tests", "cpp", "mock_qrkit"),
                    "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_eigen_adapter.cpp"),
                    "-o", exe, lib, f"-Wl,-rpath,{os.path.dirname(lib)}"], check=True)
    return exe


def test_adapter_compiles_against_the_mock_and_refuses_without_gpu(tmp_path):
    from qrkit_b200 import capi
    exe = _build(tmp_path)
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present; the gpu test runs the program")
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 77, res.stdout + res.stderr


def test_adapter_is_inert_without_eigen(tmp_path):
    """Without <Eigen/Sparse> on the include path the header defines nothing and still compiles."""
    src = tmp_path / "noeigen.cpp"
    src.write_text('#include "qrkit_b200/EigenAdapter.hpp"\n#ifdef QRKIT_B200_HAVE_EIGEN\n#error unexpected\n#endif\nint main() { return 0; }\n')
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-fsyntax-only", str(src)], check=True)


@pytest.mark.gpu
def test_adapter_solves_on_gpu(tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "All passed." in res.stdout
