"""Config 3 as the reference calls it: compute(J) then solve(b), on device buffers (development measurement)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench_extra as be
from qrkit_b200 import capi
from qrkit_b200.capi import QRK_DEVICE, QrkDesc, check
L = capi.lib()
n = 1_000_000
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    stream = C.c_void_p(s.cuda_stream)
    J1, J2, rhs = be.ellipse_device(n)
    x = torch.empty(n + 5, dtype=torch.float64, device="cuda")
    d = QrkDesc()
    d.kind, d.num_blocks, d.block_rows, d.block_cols, d.pivoting, d.border_cols = capi.QRK_BLOCK_ANGULAR, n, 2, 1, 0, 5
    h = C.c_void_p()
    check(L.qrk_create(C.byref(d), C.byref(h)))
    check(L.qrk_set_stream(h, stream), h)
    check(L.qrk_set_border(h, be.vp(J2), 2 * n, QRK_DEVICE), h)
    comp = lambda: check(L.qrk_compute(h, be.vp(J1), QRK_DEVICE), h)
    solve = lambda: check(L.qrk_solve(h, be.vp(rhs), 2 * n, be.vp(x), n + 5, 1, QRK_DEVICE), h)
    both = lambda: (comp(), solve())
    print("compute %.2f us  solve %.2f us  compute+solve %.2f us  (graphs %s)" % (
        be.time_steps(comp, 40, 5) * 1e3, be.time_steps(solve, 40, 5) * 1e3, be.time_steps(both, 40, 5) * 1e3,
        "off" if os.environ.get("QRK_NO_GRAPH") else "on"))
