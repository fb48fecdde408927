#!/usr/bin/env python
"""Development helper: link a VARIANT of libqrkit_b200.so in which only some translation units are recompiled with extra
-D flags (A/B measurements of tuning parameters on the GPU box in one gpurun call).
usage: python tools/build_variant.py <out.so> <tu-prefix> <flags...>     e.g.  gpurun_out/lib_u4.so angular -DQRK_ANG_U1=4
Select the variant at run time with QRKIT_B200_LIB=<out.so> (qrkit_b200/capi.py)."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qrkit_b200 import build as B

out, prefix, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build_library()                                    # the default objects exist and are current
obj_dir = os.path.join(B._PKG, "build")
var_dir = os.path.join(obj_dir, "variant_" + os.path.basename(out).replace(".", "_"))
os.makedirs(var_dir, exist_ok=True)
compile_flags = [f for f in B.NVCC_FLAGS if f != "-shared"]
objs, jobs = [], []
for name, src, extra in B.translation_units():
    if name.startswith(prefix):
        o = os.path.join(var_dir, name)
        jobs.append([B._nvcc(), *compile_flags, *extra, *flags, "-I", B.INCLUDE, "-c", "-o", o, src])
        objs.append(o)
    else:
        objs.append(os.path.join(obj_dir, name))
with ThreadPoolExecutor(max_workers=8) as ex:
    for r in ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs):
        if r.returncode != 0:
            raise SystemExit(r.stdout + r.stderr)
subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC", "-o", out, *objs], check=True)
print(out)
