"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/qrkit_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback).
No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from qrkit_b200.build import build_library
    build_library()
    from qrkit_b200 import capi as m
    return m


def test_library_exports_every_declared_symbol(capi):
    names = capi.declared_symbols()
    assert len(names) >= 30 and len(set(names)) == len(names)
    L = capi.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # and nothing but the C ABI (plus weak C++ template instantiations) is exported
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    strong = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    assert all(s.startswith("qrk_") for s in strong), [s for s in strong if not s.startswith("qrk_")]
    assert sorted(strong) == sorted(names)


def test_header_cites_the_reference_for_each_entry_point(capi):
    text = open(capi.HEADER).read()
    assert len(re.findall(r"\w+\.h:\d+", text)) >= 20


def test_header_is_plain_c(capi, tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "qrkit_b200.h"\nint main(void){ qrk_desc_t d; (void)d; return 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_status_strings_and_version(capi):
    L = capi.lib()
    assert L.qrk_version() >= 100
    assert b"no CPU fallback" in L.qrk_status_string(capi.QRK_STATUS_NO_DEVICE)
    assert L.qrk_status_string(0) == b"ok"


def test_no_device_means_no_compute(capi):
    """Without a GPU the library must refuse, not fall back to the CPU."""
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    d = capi.QrkDesc()
    d.kind, d.num_blocks, d.block_rows, d.block_cols = capi.QRK_BLOCK_DIAGONAL, 4, 8, 4
    h = C.c_void_p()
    assert capi.lib().qrk_create(C.byref(d), C.byref(h)) == capi.QRK_STATUS_NO_DEVICE
    assert not h
    import qrkit_b200 as qk
    import numpy as np
    with pytest.raises(qk.QrkError):
        qk.BlockDiagonalSparseQR(qk.SparseBlockDiagonal(np.ones(32), block_rows=8, block_cols=4))


def test_product_does_not_touch_the_oracle():
    """The product package must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "qrkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "qrkit_oracle" not in text, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read()


def test_ctypes_descriptor_matches_the_c_struct(capi, tmp_path):
    """qrk_desc_t as the Python binding declares it must be byte-compatible with the header (size and every field offset)."""
    import ctypes as C
    fields = [f[0] for f in capi.QrkDesc._fields_ if f[0] != "reserved"]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "qrkit_b200.h"', 'int main(void){', '  printf("%zu\\n", sizeof(qrk_desc_t));']
    for f in fields:
        lines.append(f'  printf("{f} %zu\\n", offsetof(qrk_desc_t, {f}));')
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.dirname(capi.HEADER), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert int(out[0]) == C.sizeof(capi.QrkDesc)
    for ln in out[1:]:
        if ln.strip():
            name, off = ln.split()
            assert getattr(capi.QrkDesc, name).offset == int(off), name


def test_new_entry_points_validate_arguments_and_refuse_without_a_device(capi):
    """The multi-GPU exchange, IPC and device-side assembly entry points: null arguments are rejected, and nothing computes
    without a CUDA device (no CPU fallback)."""
    import ctypes as C
    L = capi.lib()
    null = C.c_void_p()
    assert L.qrk_angular_xchg_buffer(null, None, None) == 1            # QRK_STATUS_INVALID_ARGUMENT
    assert L.qrk_angular_p2p_attach(null, None, 2, 0) == 1
    assert L.qrk_angular_p2p_status(null, None) == 1
    assert L.qrk_ipc_export(null, null) == 1 and L.qrk_ipc_close(null) == 1
    buf = (C.c_double * 16)()
    p = C.cast(buf, C.c_void_p)
    assert L.qrk_ellipse_points(null, null, 8, 1.0, 1.0, 0.0, 0.0, 0.0, None) == 1
    assert L.qrk_ellipse_assemble(null, null, null, 8, null, null, null, null, None) == 1
    if capi.device_count() == 0:
        assert L.qrk_ellipse_points(p, p, 8, 1.0, 1.0, 0.0, 0.0, 0.0, None) == 4      # QRK_STATUS_NO_DEVICE
        assert L.qrk_ellipse_assemble(p, p, p, 8, p, p, p, None, None) == 4
