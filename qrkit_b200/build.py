"""In-tree build of the CUDA library (sm_100a only): nvcc -> qrkit_b200/lib/libqrkit_b200.so.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles
without a GPU, so this runs in the build container (``__graft_entry__.build()``)."""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libqrkit_b200.so")
INCLUDE = os.path.join(os.path.dirname(_PKG), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only: no PTX for other targets, no fallback arch
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-cudart", "static",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: qrkit_b200 has no CPU fallback and cannot be built without the CUDA toolkit")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)
                                                                 if f.endswith(".h")]
    return any(os.path.getmtime(d) > t for d in deps)


ANGULAR_M2 = range(1, 9)      # border widths instantiated for the block-angular kernels (angular_inst.cu)


def translation_units():
    """(object name, source, extra flags): capi.cu once, angular_inst.cu once per border width."""
    tus = [("capi.o", os.path.join(CSRC, "capi.cu"), []), ("banded.o", os.path.join(CSRC, "banded_inst.cu"), []),
           ("structure.o", os.path.join(CSRC, "structure.cpp"), [])]
    tus += [(f"angular_m{k}.o", os.path.join(CSRC, "angular_inst.cu"), [f"-DQRK_M2={k}"]) for k in ANGULAR_M2]
    return tus


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu (in parallel, one nvcc per translation unit) and link one shared library exporting
    the C ABI of include/qrkit_b200.h."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(_PKG, "build")
    os.makedirs(obj_dir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f not in ("-shared",)]
    compile_flags += os.environ.get("QRK_NVCC_EXTRA", "").split()      # development switch, e.g. -DQRK_WY_ROWPANEL (A/B builds)

    def compile_one(tu):
        name, src, extra = tu
        obj = os.path.join(obj_dir, name)
        cmd = [_nvcc(), *compile_flags, *extra, "-I", INCLUDE, "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, translation_units()))
    if verbose:
        for _, err in results:
            print(err)
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC",
            "-o", LIB_PATH, *[o for o, _ in results]]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(link) + "\n" + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
