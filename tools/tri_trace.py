import ctypes as C, os, sys, subprocess, json
os.environ["QRKIT_B200_LIB"] = os.path.abspath("tools/variants/tritrace.so")
sys.path.insert(0, ".")
import bench_extra
from qrkit_b200 import capi
import torch
L = capi.lib()
class A: pass
a = A(); a.steps = 8; a.warmup = 1; a.no_cpu = True
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    line = bench_extra.bench_angular_wide(a, L, C.c_void_p(s.cuda_stream))
print({k: round(v["ms_per_step"], 3) for k, v in line["right_solver"].items()})
out = (C.c_longlong * 16)()
L.qrk_debug_tri_trace.argtypes = [C.POINTER(C.c_longlong)]
print("rc", L.qrk_debug_tri_trace(out))
t = list(out)
names = ["step start", "local candidates", "cluster.sync 1", "global max over DSMEM", "owner: swap, reflector, push", "cluster.sync 2", "update + downdate", "__syncthreads"]
for i in range(1, 8): print(f"{t[i]-t[i-1]:7d} cycles  {names[i]}")
print("total", t[7]-t[0])
L.qrk_debug_panel_trace.argtypes = [C.POINTER(C.c_longlong)]
print("panel rc", L.qrk_debug_panel_trace(out))
t = list(out)
pn = ["start", "first cluster.sync", "panel loaded", "column 0", "columns 1..7", "G = V^T V reduction", "T factor", "panel stored", "last cluster.sync"]
for i in range(1, 9): print(f"{t[i]-t[i-1]:7d} cycles  {pn[i]}")
print("panel total", t[8]-t[0])
