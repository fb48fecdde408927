#!/usr/bin/env python
"""bench_extra.py — device-timed measurements of the other BASELINE configs (3: block-angular ellipse Jacobian at
1M points, 5: mixed 32x16..128x64 blocks, 2 as two API calls).  One JSON line per workload; the headline line
the driver reads is bench.py's.  Inputs are generated on the device; every call goes through the C ABI."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

from bench import measured_peaks
from qrkit_b200 import capi
from qrkit_b200.capi import QRK_DEVICE, QrkDesc, check

SEED_A = 0x51524B49
# 64 FP64 FMA per clock per SM (DFMA and DMMA alike, measured: profiles/r01_fp64_pipes_b200.txt) x 148 SMs x 1.965 GHz max clock
FP64_NOMINAL_TFLOPS = 2 * 64 * 148 * 1.965e9 / 1e12


def vp(t):
    return C.c_void_p(t.data_ptr())


def time_steps(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _cpu(fn, rows, sample, cores=1):
    """Time the CPU oracle (restatement of the reference's algorithm, oracle/) once on a bounded sample."""
    import time
    t0 = time.perf_counter()
    fn()
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": "rows/s", "cores": cores, "kind": "port", "seconds": dt, "sample": sample}


def cpu_angular(n):
    from helpers import ellipse_problem
    from oracle import oracle as orc
    J1, J2, rhs = ellipse_problem(n)
    def run():
        ref = orc.BlockAngularOracle(J2, br=np.full(n, 2, dtype=np.int32), bc=np.full(n, 1, dtype=np.int32), values=J1, left_colpiv=True, right_kind=0)
        ref.solve(rhs)
    return _cpu(run, 2 * n, f"ellipse Jacobian at N={n} points, BlockAngular<BlockDiagonal 2x1 ColPiv, dense ColPiv> factorize + solve, reference-faithful (explicit sparse Q1, dense tall ColPiv QR), 1 thread")


def cpu_mixed(nb):
    from helpers import uniform_blocks, vector
    from oracle import oracle as orc
    br, bc = mixed_sizes(nb)
    vals = np.concatenate([uniform_blocks(1, int(r), int(c), block0=i) for i, (r, c) in enumerate(zip(br, bc))])
    b = vector(int(br.sum()), seed=17)
    def run():
        ref = orc.BlockDiagonalOracle(br, bc, vals, colpiv=False)
        ref.solve(b)
    return _cpu(run, int(br.sum()), f"{nb} mixed blocks 32x16..128x64, BlockDiagonalSparseQR<HouseholderQR> factorize + solve, reference-faithful (explicit Q_i, sparse Q/R assembly), 1 thread")


def cpu_angular_wide(nb, r, c, m2):
    """Reference test 4 (test/test-qrkit.cpp:283-327): BlockAngular<BlockDiagonal 7x2 ColPiv, dense ColPiv>."""
    from helpers import uniform_blocks, vector
    from oracle import oracle as orc
    n = nb * r
    vals = uniform_blocks(nb, r, c)
    J2 = np.asfortranarray(vector(m2 * n, SEED_A + 7, 0.5, 5.0).reshape(m2, n).T)
    b = vector(n, seed=SEED_A + 5)
    def run():
        ref = orc.BlockAngularOracle(J2, br=np.full(nb, r, dtype=np.int32), bc=np.full(nb, c, dtype=np.int32), values=vals, left_colpiv=True, right_kind=0)
        ref.solve(b)
    return _cpu(run, n, f"reference test 4 at its own sizes ({nb} blocks {r}x{c} + dense {n}x{m2} border), the whole problem, reference-faithful (explicit sparse Q1, dense tall ColPiv QR), 1 thread")


def cpu_banded(nb):
    import scipy.sparse as sp
    from helpers import reference_style_windows, uniform_blocks, vector
    from oracle import oracle as orc
    br, bc, ov = 16, 24, 16
    slabs = uniform_blocks(nb, br, bc).reshape(nb, bc, br)
    jj, ii = np.meshgrid(np.arange(bc), np.arange(br), indexing="ij")
    rows = (np.arange(nb)[:, None, None] * br + ii[None]).reshape(-1)
    cols = (np.arange(nb)[:, None, None] * (bc - ov) + jj[None]).reshape(-1)
    A = sp.csc_matrix((slabs.reshape(-1), (rows, cols)), shape=(nb * br, (nb - 1) * (bc - ov) + bc))
    blocks = reference_style_windows(nb, br, bc, ov, 2)
    b = vector(nb * br, seed=3)
    def run():
        ref = orc.BandedOracle(A, blocks)
        ref.solve(b)
    return _cpu(run, nb * br, f"{nb} block rows 16x24 step 8, BandedBlockedSparseQR with the reference's merged windows ({len(blocks)} of them) factorize + solve, 1 thread")


def ellipse_device(n):
    """Ellipse-fit Jacobian at the initial LM iterate (bench/bench_sparse_qr_extra.cpp:79-114, 221-282) on the device."""
    a, b, x0, y0, r = 7.5, 2.0, 17.0, 23.0, 0.23
    t = torch.arange(n, dtype=torch.float64, device="cuda") * (1.3 * math.pi / n)
    px = x0 + a * torch.cos(t) * math.cos(r) - b * torch.sin(t) * math.sin(r)
    py = y0 + a * torch.cos(t) * math.sin(r) + b * torch.sin(t) * math.cos(r)
    pa = 0.5 * (px.max() - px.min()); pb = 0.5 * (py.max() - py.min())
    cx = 0.5 * (px.max() + px.min()); cy = 0.5 * (py.max() + py.min())
    ct, st, cr, sr = torch.cos(t), torch.sin(t), 1.0, 0.0
    J1 = torch.empty(n, 2, dtype=torch.float64, device="cuda")
    J1[:, 0] = pa * cr * st + pb * sr * ct
    J1[:, 1] = pa * sr * st - pb * cr * ct
    J2 = torch.zeros(5, 2 * n, dtype=torch.float64, device="cuda")      # column-major 2n x 5
    J2[0, 0::2] = -ct * cr; J2[1, 0::2] = st * sr; J2[2, 0::2] = -1; J2[4, 0::2] = pa * ct * sr + pb * st * cr
    J2[0, 1::2] = -ct * sr; J2[1, 1::2] = -st * cr; J2[3, 1::2] = -1; J2[4, 1::2] = -pa * ct * cr + pb * st * sr
    rhs = torch.empty(2 * n, dtype=torch.float64, device="cuda")
    rhs[0::2] = px - (pa * ct * cr - pb * st * sr + cx)
    rhs[1::2] = py - (pa * ct * sr + pb * st * cr + cy)
    return J1.reshape(-1).contiguous(), J2, rhs


def bench_angular(args, L, stream):
    n = args.points
    J1, J2, rhs = ellipse_device(n)
    x = torch.empty(n + 5, dtype=torch.float64, device="cuda")
    out = {}
    graph_ms = None
    for piv in (0, 1):
        d = QrkDesc()
        d.kind, d.num_blocks, d.block_rows, d.block_cols, d.pivoting, d.border_cols = capi.QRK_BLOCK_ANGULAR, n, 2, 1, piv, 5
        h = C.c_void_p()
        check(L.qrk_create(C.byref(d), C.byref(h)))
        check(L.qrk_set_stream(h, stream), h)
        check(L.qrk_set_border(h, vp(J2), 2 * n, QRK_DEVICE), h)

        def step():
            check(L.qrk_compute_solve(h, vp(J1), vp(rhs), vp(x), QRK_DEVICE), h)
        l0 = C.c_int64(); L.qrk_launch_count(h, C.byref(l0))
        ms = time_steps(step, args.steps, args.warmup)
        l1 = C.c_int64(); L.qrk_launch_count(h, C.byref(l1))
        out[piv] = (ms, (l1.value - l0.value) // (args.steps + args.warmup))
        if piv == 0 and getattr(args, "graphs", True):
            # the same step replayed from a CUDA graph (the library's calls are stream-ordered and capture-safe): what the three
            # launches cost when their dependencies are resolved on the device instead of through the stream
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
                    step()
                graph_ms = time_steps(g.replay, args.steps, args.warmup)
            except Exception as e:
                print(json.dumps({"graph_capture_failed": str(e)[:200]}), file=sys.stderr, flush=True)
        L.qrk_destroy(h)
    peak, src = measured_peaks()
    ms, launches = out[0]
    bytes_per_point = 248          # J1 16 + J2 80 + b 16 in; packed 16(+tau 8) ... second pass; see DESIGN.md
    achieved = bytes_per_point * n / (ms * 1e-3) / 1e9
    line = {"workload": f"block-angular ellipse Jacobian, N={n} points: {2*n} x ({n}+5), fused compute+solve (BASELINE config 3)",
            "metric": "rows/s", "value": 2 * n / (ms * 1e-3), "ms_per_step": ms, "launches_per_step": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes_per_point": bytes_per_point, "peak_source": src},
            "colpiv_left": {"ms_per_step": out[1][0], "value": 2 * n / (out[1][0] * 1e-3)},
            "steps": args.steps, "warmup": args.warmup, "dtype": "f64"}
    if graph_ms is not None:
        line["cuda_graph_replay"] = {"ms_per_step": graph_ms, "value": 2 * n / (graph_ms * 1e-3),
                                     "roofline_frac": bytes_per_point * n / (graph_ms * 1e-3) / 1e9 / peak}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_angular(min(n, args.cpu_points))
    return line


def mixed_sizes(nb, seed=SEED_A):
    from helpers import splitmix64
    hsh = splitmix64(np.arange(nb, dtype=np.uint64) ^ np.uint64(seed))
    br = (32 + 16 * (hsh % np.uint64(7)).astype(np.int64)).astype(np.int32)
    return br, (br // 2).astype(np.int32)


def bench_mixed(args, L, stream):
    nb = args.mixed_blocks
    br, bc = mixed_sizes(nb)
    total = int((br.astype(np.int64) * bc).sum())
    rows, cols = int(br.sum()), int(bc.sum())
    A = torch.empty(total, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(A), SEED_A, 0, total, 1, 0, 0.5, 5.0, stream))
    b = torch.empty(rows, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(b), SEED_A + 5, 0, rows, 1, 0, -1.0, 1.0, stream))
    x = torch.empty(cols, dtype=torch.float64, device="cuda")
    res = {}
    for piv in (0, 1):
        d = QrkDesc()
        d.kind, d.num_blocks, d.pivoting = 0, nb, piv
        d.rows = br.ctypes.data_as(C.POINTER(C.c_int32)); d.cols = bc.ctypes.data_as(C.POINTER(C.c_int32))
        h = C.c_void_p()
        check(L.qrk_create(C.byref(d), C.byref(h)))
        check(L.qrk_set_stream(h, stream), h)

        def step():
            check(L.qrk_compute_solve(h, vp(A), vp(b), vp(x), QRK_DEVICE), h)
        res[piv] = time_steps(step, args.steps, args.warmup)
        L.qrk_destroy(h)
    # the two-call path on the stored factors: solve(b) and matrixQ().transpose() * b (bd_generic_op_kernel)
    d = QrkDesc()
    d.kind, d.num_blocks, d.pivoting = 0, nb, 0
    d.rows = br.ctypes.data_as(C.POINTER(C.c_int32)); d.cols = bc.ctypes.data_as(C.POINTER(C.c_int32))
    h = C.c_void_p()
    check(L.qrk_create(C.byref(d), C.byref(h)))
    check(L.qrk_set_stream(h, stream), h)
    ms_f = time_steps(lambda: check(L.qrk_compute(h, vp(A), QRK_DEVICE), h), args.steps, args.warmup)
    ms_s = time_steps(lambda: check(L.qrk_solve(h, vp(b), rows, vp(x), cols, 1, QRK_DEVICE), h), args.steps, args.warmup)
    y = torch.empty_like(b)
    ms_q = time_steps(lambda: check(L.qrk_apply_qt(h, vp(b), rows, vp(y), rows, 1, QRK_DEVICE), h), args.steps, args.warmup)
    L.qrk_destroy(h)
    peak, src = measured_peaks()
    r64, c64 = br.astype(np.float64), bc.astype(np.float64)
    alg_bytes = float((16 * r64 * c64 + 8 * r64 + 16 * c64).sum())
    op_bytes = float((8 * r64 * c64 + 16 * r64 + 8 * c64).sum())
    flops = float((2 * r64 * c64 * c64 - (2.0 / 3.0) * c64 ** 3 + 4 * r64 * c64 + c64 * c64).sum())
    ms = res[0]
    line = {"workload": f"mixed block-diagonal, {nb} blocks 32x16..128x64 (BASELINE config 5), fused QR+solve",
            "metric": "rows/s", "value": rows / (ms * 1e-3), "ms_per_step": ms, "rows": rows,
            "roofline": {"hbm": {"achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak},
                         "fp64": {"achieved": flops / (ms * 1e-3) / 1e12, "peak": FP64_NOMINAL_TFLOPS, "unit": "TFLOP/s (nominal FP64: 64 FMA/clk/SM at 1.965 GHz)",
                                  "frac": flops / (ms * 1e-3) / 1e12 / FP64_NOMINAL_TFLOPS},
                         "algorithmic_bytes": alg_bytes, "flops": flops, "peak_source": src},
            "colpiv": {"ms_per_step": res[1], "value": rows / (res[1] * 1e-3)},
            "two_call": {"compute_ms": ms_f, "solve_ms": ms_s, "solve_hbm_frac": op_bytes / (ms_s * 1e-3) / 1e9 / peak,
                         "apply_qt_ms": ms_q, "apply_qt_hbm_frac": op_bytes / (ms_q * 1e-3) / 1e9 / peak,
                         "note": "solve / Q^T b on stored factors read packed V (8rc) + tau + the vector once"},
            "steps": args.steps, "dtype": "f64"}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_mixed(min(nb, args.cpu_mixed_blocks))
    return line


def bench_classes(args, L, stream):
    """Config 5 one size class at a time (uniform r x r/2 blocks, nb/7 each): where the mixed time goes."""
    peak, src = measured_peaks()
    per = []
    shapes = [(32 + 16 * k, 16 + 8 * k) for k in range(7)] if not args.shapes else \
        [tuple(int(v) for v in a.split("x")) for a in args.shapes.split(",")]
    for (r, c) in shapes:
        nb = args.class_blocks or args.mixed_blocks // 7
        A = torch.empty(nb * r * c, dtype=torch.float64, device="cuda")
        check(L.qrk_synth_fill(vp(A), SEED_A, 0, nb, r, c, 0.5, 5.0, stream))
        b = torch.empty(nb * r, dtype=torch.float64, device="cuda")
        check(L.qrk_synth_fill(vp(b), SEED_A + 5, 0, nb * r, 1, 0, -1.0, 1.0, stream))
        x = torch.empty(nb * c, dtype=torch.float64, device="cuda")
        d = QrkDesc()
        d.kind, d.num_blocks, d.block_rows, d.block_cols, d.pivoting = 0, nb, r, c, int(getattr(args, 'class_piv', 0))
        h = C.c_void_p()
        check(L.qrk_create(C.byref(d), C.byref(h)))
        check(L.qrk_set_stream(h, stream), h)
        ms = time_steps(lambda: check(L.qrk_compute_solve(h, vp(A), vp(b), vp(x), QRK_DEVICE), h), args.steps, args.warmup)
        L.qrk_destroy(h)
        alg = nb * (16.0 * r * c + 8 * r + 16 * c)
        fl = nb * (2.0 * r * c * c - (2.0 / 3.0) * c ** 3 + 4 * r * c + c * c)
        per.append({"block": [r, c], "blocks": nb, "ms": round(ms, 4), "us_per_block_per_sm": round(ms * 1e3 * 148 / nb, 3),
                    "hbm_frac": round(alg / (ms * 1e-3) / 1e9 / peak, 4), "fp64_frac": round(fl / (ms * 1e-3) / 1e12 / FP64_NOMINAL_TFLOPS, 4)})
    return {"workload": "config 5 per size class (uniform blocks, fused QR+solve, unpivoted)", "classes": per,
            "total_ms": sum(p["ms"] for p in per), "peak_source": src}


def bench_banded(args, L, stream):
    """BASELINE config 4: block-banded, nb block rows of 16x24, column step 8 (overlap 16), single GPU."""
    nb, br, bc, ov = args.banded_blocks, 16, 24, 16
    n_rows, n_cols = nb * br, (nb - 1) * (bc - ov) + bc
    A = torch.empty(nb * br * bc, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(A), SEED_A, 0, nb, br, bc, 0.5, 5.0, stream))
    b = torch.empty(n_rows, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(b), SEED_A + 5, 0, n_rows, 1, 0, -1.0, 1.0, stream))
    x = torch.empty(n_cols, dtype=torch.float64, device="cuda")
    d = QrkDesc()
    d.kind, d.num_blocks, d.block_rows, d.block_cols, d.block_overlap = capi.QRK_BANDED_BLOCKED, nb, br, bc, ov
    h = C.c_void_p()
    check(L.qrk_create(C.byref(d), C.byref(h)))
    check(L.qrk_set_stream(h, stream), h)
    ms = time_steps(lambda: check(L.qrk_compute_solve(h, vp(A), vp(b), vp(x), QRK_DEVICE), h), max(2, args.steps // 4), 1)
    L.qrk_destroy(h)
    peak, src = measured_peaks()
    # 8 (nnz A + nnz R + nnz V) + 8 (rows + 2 cols) bytes
    alg = 8.0 * (nb * br * bc + n_cols * bc + nb * br * bc) + 8.0 * (n_rows + 2 * n_cols)
    line = {"workload": f"block-banded, {nb} block rows of {br}x{bc}, step {bc-ov} (BASELINE config 4), fused QR+solve, single GPU",
            "metric": "rows/s", "value": n_rows / (ms * 1e-3), "ms_per_step": ms, "rows": n_rows, "cols": n_cols,
            "us_per_window": ms * 1e3 / nb,
            "roofline": {"bound": "latency (sequential window chain); HBM fraction for the record", "achieved": alg / (ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak, "peak_source": src},
            "dtype": "f64"}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_banded(min(nb, args.cpu_banded_blocks))
    return line


def bench_angular_wide(args, L, stream):
    """Reference test 4 / 5 sizes (test/test-qrkit.cpp:388-391): 1024 left blocks of 7x2 + dense 7168 x 384 border; right solver
    ColPivHouseholderQR (test 4) and unpivoted BlockedThinDenseQR (test 5)."""
    nb, r, c, m2 = 1024, 7, 2, 384
    n = nb * r
    A = torch.empty(nb * r * c, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(A), SEED_A, 0, nb, r, c, 0.5, 5.0, stream))
    J2 = torch.empty(m2 * n, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(J2), SEED_A + 7, 0, m2 * n, 1, 0, 0.5, 5.0, stream))
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(b), SEED_A + 5, 0, n, 1, 0, -1.0, 1.0, stream))
    x = torch.empty(nb * c + m2, dtype=torch.float64, device="cuda")
    flops = 2.0 * (n - nb * c) * m2 * m2 - (2.0 / 3.0) * m2 ** 3
    steps = max(2, args.steps // 4)
    res = {}
    for name, right, left in (("colpiv", 0, 0), ("unpivoted", 1, 0), ("banded_left_colpiv", 0, 1)):
        d = QrkDesc()
        d.kind, d.num_blocks, d.block_rows, d.block_cols, d.pivoting, d.border_cols = capi.QRK_BLOCK_ANGULAR, nb, r, c, 1, m2
        d.right_solver = right
        if left:            # LeftSolver = BandedBlockedSparseQR (test/test-qrkit.cpp:44-48): the same 7x2 blocks as slabs, overlap 0
            d.left_solver, d.block_overlap, d.pivoting = 1, 0, 0
        h = C.c_void_p()
        check(L.qrk_create(C.byref(d), C.byref(h)))
        check(L.qrk_set_stream(h, stream), h)
        check(L.qrk_set_border(h, vp(J2), n, QRK_DEVICE), h)
        l0 = C.c_int64(); L.qrk_launch_count(h, C.byref(l0))
        ms = time_steps(lambda: check(L.qrk_compute_solve(h, vp(A), vp(b), vp(x), QRK_DEVICE), h), steps, 3)      # (3 warm-ups: eager, capture, first replay)
        l1 = C.c_int64(); L.qrk_launch_count(h, C.byref(l1))
        res[name] = {"ms_per_step": ms, "value": n / (ms * 1e-3), "launches_per_step": (l1.value - l0.value) // (steps + 3),
                     "border_qr_gflops": flops / (ms * 1e-3) / 1e9}
        if getattr(args, "graphs", True) and name != "banded_left_colpiv":
            # ~245 launches on two streams per step: the same step replayed from a CUDA graph (fork / join of the look-ahead
            # stream captured with it)
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
                    check(L.qrk_compute_solve(h, vp(A), vp(b), vp(x), QRK_DEVICE), h)
                res[name]["cuda_graph_replay_ms_per_step"] = time_steps(g.replay, steps, 1)
                del g
            except Exception as e:
                res[name]["cuda_graph_replay_ms_per_step"] = None
                print(json.dumps({"wide_graph_capture_failed": name, "error": str(e)[:200]}), file=sys.stderr, flush=True)
                torch.cuda.synchronize()
        L.qrk_destroy(h)
    line = {"workload": f"block-angular, reference test 4/5 sizes: {nb} blocks {r}x{c} + dense {n}x{m2} border; blocked compact-WY (DMMA) first stage, ColPiv on the triangle in one cluster launch",
            "metric": "rows/s", "value": res["colpiv"]["value"], "ms_per_step": res["colpiv"]["ms_per_step"],
            "right_solver": res, "dtype": "f64",
            "roofline": {"bound": "fp64", "achieved": res["colpiv"]["border_qr_gflops"] / 1e3, "peak": FP64_NOMINAL_TFLOPS, "unit": "TFLOP/s",
                         "frac": res["colpiv"]["border_qr_gflops"] / 1e3 / FP64_NOMINAL_TFLOPS, "traffic": None,
                         "peak_source": "nominal FP64 (64 FMA/clk/SM x 148 SMs x 1.965 GHz; DFMA = DMMA peak, profiles/r01_fp64_pipes_b200.txt)",
                         "flops": flops}}
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_angular_wide(nb, r, c, m2)
    return line


def bench_two_call(args, L, stream):
    nb, r, c = 1_000_000, 8, 4
    A = torch.empty(nb * r * c, dtype=torch.float64, device="cuda")
    b = torch.empty(nb * r, dtype=torch.float64, device="cuda")
    x = torch.empty(nb * c, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(A), SEED_A, 0, nb, r, c, 0.5, 5.0, stream))
    check(L.qrk_synth_fill(vp(b), SEED_A + 5, 0, nb, r, 0, -1.0, 1.0, stream))
    d = QrkDesc()
    d.kind, d.num_blocks, d.block_rows, d.block_cols, d.pivoting = 0, nb, r, c, 0
    h = C.c_void_p()
    check(L.qrk_create(C.byref(d), C.byref(h)))
    check(L.qrk_set_stream(h, stream), h)
    ms_f = time_steps(lambda: check(L.qrk_compute(h, vp(A), QRK_DEVICE), h), args.steps, args.warmup)
    ms_s = time_steps(lambda: check(L.qrk_solve(h, vp(b), nb * r, vp(x), nb * c, 1, QRK_DEVICE), h), args.steps, args.warmup)
    y = torch.empty_like(b)
    ms_q = time_steps(lambda: check(L.qrk_apply_qt(h, vp(b), nb * r, vp(y), nb * r, 1, QRK_DEVICE), h), args.steps, args.warmup)
    L.qrk_destroy(h)
    peak, _ = measured_peaks()
    line = {"workload": "config 2 as separate API calls: compute(), solve(b), matrixQ().transpose()*b",
            "factor": {"ms": ms_f, "bytes_per_block": 544, "frac": 544 * nb / (ms_f * 1e-3) / 1e9 / peak},
            "solve": {"ms": ms_s, "bytes_per_block": 384, "frac": 384 * nb / (ms_s * 1e-3) / 1e9 / peak},
            "apply_qt": {"ms": ms_q, "bytes_per_block": 256 + 32 + 128, "frac": 416 * nb / (ms_q * 1e-3) / 1e9 / peak}}
    return line


# ------------------------------------------------------------------------------------------------------------------
# Multi-GPU (one process per GPU under torchrun).  The NAMED shape is sharded over the ranks ("strong"), or
# every rank gets the named shape ("weak", --scaling weak).  Timing: barrier + CUDA events per rank, max over ranks.
# ------------------------------------------------------------------------------------------------------------------
def _dist_time(fn, steps, warmup):
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    return float(t.item())


def bench_angular_dist(args, L, stream):
    """Config 3 on G GPUs: contiguous point ranges, per-GPU TSQR triangle, ONE NCCL all-gather of G x 20 doubles, redundant
    root on every rank, local back substitution (x1 stays sharded, x2 replicated)."""
    import torch.distributed as dist
    from qrkit_b200.distributed import block_range
    G, rank = dist.get_world_size(), dist.get_rank()
    n_total = args.points * (G if args.scaling == "weak" else 1)
    lo, hi = block_range(n_total, G, rank)
    n = hi - lo
    J1f, J2f, rhsf = ellipse_device(n_total)            # every rank builds the same global problem, then keeps its slice
    J1 = J1f[2 * lo:2 * hi].contiguous()
    J2 = J2f[:, 2 * lo:2 * hi].contiguous()
    rhs = rhsf[2 * lo:2 * hi].contiguous()
    del J1f, J2f, rhsf
    x = torch.empty(n + 5, dtype=torch.float64, device="cuda")
    d = QrkDesc()
    d.kind, d.num_blocks, d.block_rows, d.block_cols, d.pivoting, d.border_cols = capi.QRK_BLOCK_ANGULAR, n, 2, 1, 0, 5
    d.device = torch.cuda.current_device()
    h = C.c_void_p()
    check(L.qrk_create(C.byref(d), C.byref(h)))
    check(L.qrk_set_stream(h, stream), h)
    check(L.qrk_angular_set_world(h, max(G, 2)), h)      # G = 1 still goes through the exchange API (one triangle)
    check(L.qrk_set_border(h, vp(J2), 2 * n, QRK_DEVICE), h)
    tsz = C.c_int64(); check(L.qrk_angular_triangle_size(h, C.byref(tsz)), h)
    tri = torch.empty(tsz.value, dtype=torch.float64, device="cuda")
    tris = torch.empty(G * tsz.value, dtype=torch.float64, device="cuda")

    def step():
        check(L.qrk_compute_solve(h, vp(J1), vp(rhs), vp(x), QRK_DEVICE), h)
        check(L.qrk_angular_local_triangle(h, vp(tri), QRK_DEVICE), h)
        dist.all_gather_into_tensor(tris, tri)            # NCCL on torch's current stream == the handle's stream
        check(L.qrk_angular_merge(h, vp(tris), G, QRK_DEVICE), h)
    ms = _dist_time(step, args.steps, args.warmup)
    # the same step replayed from a CUDA graph (kernels + the NCCL all-gather captured once): at 125k points per GPU the
    # eager loop is bound by the host issuing ~8 calls per step, not by the device
    ms_graph = None
    try:
        if not getattr(args, "graphs", True):
            raise RuntimeError("graphs disabled")
        torch.cuda.synchronize(); dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
            step()
        ms_graph = _dist_time(g.replay, args.steps, args.warmup)
    except Exception as e:                                   # capture is an optimisation of the measurement, not of the product
        ms_graph = None
        if rank == 0 and getattr(args, "graphs", True):
            print(json.dumps({"graph_capture_failed": str(e)[:200]}), file=sys.stderr, flush=True)
    # fused peer exchange: the triangles travel over NVLink inside the TSQR root kernel (qrk_angular_p2p_attach); the step is
    # ONE library call again, no NCCL and no merge launch on the data path
    ms_fused = ms_fused_graph = None
    if G > 1:
        from qrkit_b200.distributed import attach_p2p
        x_nccl = x.clone()
        imported = attach_p2p(h)

        def step_fused():
            check(L.qrk_compute_solve(h, vp(J1), vp(rhs), vp(x), QRK_DEVICE), h)
        ms_fused = _dist_time(step_fused, args.steps, args.warmup)
        to = C.c_int32(-1); check(L.qrk_angular_p2p_status(h, C.byref(to)), h)
        fused_matches = bool(torch.equal(x, x_nccl)) and to.value == 0
        try:                                   # the step counter lives on the device, so a captured step replays correctly
            if not getattr(args, "graphs", True):
                raise RuntimeError("graphs disabled")
            torch.cuda.synchronize(); dist.barrier()
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, stream=torch.cuda.current_stream()):
                step_fused()
            ms_fused_graph = _dist_time(g2.replay, args.steps, args.warmup)
            check(L.qrk_angular_p2p_status(h, C.byref(to)), h)
            fused_matches = fused_matches and bool(torch.equal(x, x_nccl)) and to.value == 0
        except Exception as e:
            ms_fused_graph = None
            if rank == 0 and getattr(args, "graphs", True):
                print(json.dumps({"fused_graph_capture_failed": str(e)[:200]}), file=sys.stderr, flush=True)
    # x2 must be bit-identical on every rank (redundant root on identical gathered triangles)
    x2 = x[n:].clone()
    x2all = torch.empty(G * 5, dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(x2all, x2)
    same = bool((x2all.view(G, 5) == x2all.view(G, 5)[0]).all().item())
    L.qrk_destroy(h)
    if G > 1:
        for pimp in imported:
            L.qrk_ipc_close(pimp)
    if rank == 0:
        peak, src = measured_peaks()
        ach = 248.0 * n_total / (ms * 1e-3) / 1e9
        return ({"workload": f"block-angular ellipse Jacobian, N={n_total} points over {G} GPU(s) (BASELINE config 3), fused compute+solve with NCCL all-gather of the per-GPU triangles",
                          "metric": "rows/s", "value": 2 * n_total / (ms * 1e-3), "ms_per_step": ms, "n_gpus": G, "scaling": args.scaling,
                          "cuda_graph_replay": None if ms_graph is None else {"ms_per_step": ms_graph, "value": 2 * n_total / (ms_graph * 1e-3)},
                          "fused_peer_exchange": None if ms_fused is None else {"ms_per_step": ms_fused, "value": 2 * n_total / (ms_fused * 1e-3), "x_identical_to_nccl_path": fused_matches,
                                                                                 "cuda_graph_replay_ms_per_step": ms_fused_graph,
                                                                                 "note": "triangles stored into the peers' buffers over NVLink inside the root kernel; no NCCL, no merge launch"},
                          "collective": {"op": "all_gather", "bytes_per_rank": tsz.value * 8, "backend": "nccl"},
                          "x2_identical_on_all_ranks": same,
                          "roofline": {"bound": "hbm", "achieved": ach, "peak": peak * G, "unit": "GB/s", "frac": ach / (peak * G), "peak_source": src},
                          "steps": args.steps, "warmup": args.warmup, "dtype": "f64"})
    return None


def bench_mixed_dist(args, L, stream):
    """Config 5 on G GPUs: contiguous block ranges balanced by bytes (not count), no collective."""
    import torch.distributed as dist
    from qrkit_b200.distributed import byte_balanced_ranges
    G, rank = dist.get_world_size(), dist.get_rank()
    nb_total = args.mixed_blocks * (G if args.scaling == "weak" else 1)
    br_all, bc_all = mixed_sizes(nb_total)
    lo, hi = byte_balanced_ranges(br_all, bc_all, G)[rank]
    br, bc = np.ascontiguousarray(br_all[lo:hi]), np.ascontiguousarray(bc_all[lo:hi])
    nb = hi - lo
    total = int((br.astype(np.int64) * bc).sum())
    rows, cols = int(br.sum()), int(bc.sum())
    A = torch.empty(total, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(A), SEED_A + rank, 0, total, 1, 0, 0.5, 5.0, stream))
    b = torch.empty(rows, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(vp(b), SEED_A + 5 + rank, 0, rows, 1, 0, -1.0, 1.0, stream))
    x = torch.empty(cols, dtype=torch.float64, device="cuda")
    d = QrkDesc()
    d.kind, d.num_blocks, d.pivoting = 0, nb, 0
    d.device = torch.cuda.current_device()
    d.rows = br.ctypes.data_as(C.POINTER(C.c_int32)); d.cols = bc.ctypes.data_as(C.POINTER(C.c_int32))
    h = C.c_void_p()
    check(L.qrk_create(C.byref(d), C.byref(h)))
    check(L.qrk_set_stream(h, stream), h)
    ms = _dist_time(lambda: check(L.qrk_compute_solve(h, vp(A), vp(b), vp(x), QRK_DEVICE), h), args.steps, args.warmup)
    L.qrk_destroy(h)
    if rank == 0:
        peak, src = measured_peaks()
        r64, c64 = br_all.astype(np.float64), bc_all.astype(np.float64)
        alg = float((16 * r64 * c64 + 8 * r64 + 16 * c64).sum())
        fl = float((2 * r64 * c64 * c64 - (2.0 / 3.0) * c64 ** 3 + 4 * r64 * c64 + c64 * c64).sum())
        rows_total = int(br_all.sum())
        return ({"workload": f"mixed block-diagonal, {nb_total} blocks 32x16..128x64 over {G} GPU(s) (BASELINE config 5), byte-balanced contiguous ranges, no collective",
                          "metric": "rows/s", "value": rows_total / (ms * 1e-3), "ms_per_step": ms, "n_gpus": G, "scaling": args.scaling,
                          "roofline": {"hbm_frac": alg / (ms * 1e-3) / 1e9 / (peak * G), "fp64_frac_nominal": fl / (ms * 1e-3) / 1e12 / (FP64_NOMINAL_TFLOPS * G), "peak_source": src},
                          "steps": args.steps, "dtype": "f64"})
    return None


def bench_lm(args, L, stream):
    """the whole ellipse fit on the device (the reference's published benchmark is the total time of such a fit,
    README.md:25-30).  Per iteration: qrk_ellipse_assemble (Jacobian blocks, border, -f, cost: one kernel) ->
    qrk_compute_solve (3 kernels) -> params += step (one axpy).  Gauss-Newton, a fixed number of iterations, no host
    synchronisation inside the loop; NOT Eigen's LevenbergMarquardt (no trust region: from the reference's own
    initialisation, bench :221-240, the undamped step converges quadratically)."""
    res = []
    for n in [int(v) for v in args.lm_points.split(",")]:
        a, b, x0, y0, r = 7.5, 2.0, 17.0, 23.0, 0.23
        px = torch.empty(n, dtype=torch.float64, device="cuda"); py = torch.empty_like(px)
        check(L.qrk_ellipse_points(vp(px), vp(py), n, a, b, x0, y0, r, stream))
        incr = 1.3 * math.pi / n
        init = torch.empty(n + 5, dtype=torch.float64, device="cuda")
        init[:n] = torch.arange(n, dtype=torch.float64, device="cuda") * incr
        init[n:] = torch.stack([0.5 * (px.max() - px.min()), 0.5 * (py.max() - py.min()), 0.5 * (px.max() + px.min()),
                                0.5 * (py.max() + py.min()), torch.zeros((), dtype=torch.float64, device="cuda")])
        params = init.clone()
        J1 = torch.empty(2 * n, dtype=torch.float64, device="cuda")
        J2 = torch.empty(5 * 2 * n, dtype=torch.float64, device="cuda")
        rhs = torch.empty(2 * n, dtype=torch.float64, device="cuda")
        step_v = torch.empty(n + 5, dtype=torch.float64, device="cuda")
        iters = args.lm_iters
        cost = torch.zeros(iters + 1, dtype=torch.float64, device="cuda")
        d = QrkDesc()
        d.kind, d.num_blocks, d.block_rows, d.block_cols, d.pivoting, d.border_cols = capi.QRK_BLOCK_ANGULAR, n, 2, 1, 1, 5
        h = C.c_void_p()
        check(L.qrk_create(C.byref(d), C.byref(h)))
        check(L.qrk_set_stream(h, stream), h)
        check(L.qrk_set_border(h, vp(J2), 2 * n, QRK_DEVICE), h)

        def fit():
            params.copy_(init)
            cost.zero_()
            for it in range(iters):
                check(L.qrk_ellipse_assemble(vp(px), vp(py), vp(params), n, vp(J1), vp(J2), vp(rhs), C.c_void_p(cost.data_ptr() + 8 * it), stream))
                check(L.qrk_compute_solve(h, vp(J1), vp(rhs), vp(step_v), QRK_DEVICE), h)
                params.add_(step_v)
            check(L.qrk_ellipse_assemble(vp(px), vp(py), vp(params), n, vp(J1), vp(J2), vp(rhs), C.c_void_p(cost.data_ptr() + 8 * iters), stream))
        ms = time_steps(fit, max(3, args.steps // 4), 2)
        L.qrk_destroy(h)
        p = params[n:].cpu().numpy()
        err = float(np.abs(p - np.array([a, b, x0, y0, r])).max())
        res.append({"points": n, "rows": 2 * n, "iterations": iters, "ms_per_fit": ms, "ms_per_iteration": ms / iters,
                    "cost": [float(v) for v in cost.cpu().numpy()], "max_param_error": err})
    published = {"QRkitBD_total_seconds": {"500": 0.005, "2000": 0.017, "10000": 0.098, "100000": 1.036, "500000": 5.342},
                 "source": "imgs/benchmark_table.png via README.md:29 (whole LM fit, unstated CPU; Eigen LevenbergMarquardt, more iterations than Gauss-Newton needs)"}
    return {"workload": "ellipse fit entirely on the device: Jacobian assembly + BlockAngularSparseQR compute+solve + update per iteration (Gauss-Newton from the reference's initialisation)",
            "fits": res, "reference_published": published, "dtype": "f64"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="angular,angular_wide,mixed,banded,two_call,lm")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--mixed-blocks", type=int, default=100_000)
    ap.add_argument("--class-blocks", type=int, default=0)
    ap.add_argument("--banded-blocks", type=int, default=100_000)
    ap.add_argument("--shapes", default="")
    ap.add_argument("--class-piv", type=int, default=0, help="classes workload: 1 = ColPivHouseholderQR per block")
    ap.add_argument("--lm-points", default="500,10000,100000,500000,1000000")
    ap.add_argument("--lm-iters", type=int, default=6)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU oracle timings (cpu_baseline)")
    ap.add_argument("--cpu-points", type=int, default=500_000, help="points of the CPU oracle sample of config 3")
    ap.add_argument("--cpu-mixed-blocks", type=int, default=5000, help="blocks of the CPU oracle sample of config 5")
    ap.add_argument("--cpu-banded-blocks", type=int, default=50_000, help="block rows of the CPU oracle sample of config 4")
    ap.add_argument("--no-graphs", dest="graphs", action="store_false", help="multi-GPU: skip the CUDA-graph replays")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="multi-GPU runs (torchrun): shard the named shape, or one named shape per GPU")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device: qrkit_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 or os.environ.get("QRK_BENCH_DIST"):
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        s = torch.cuda.Stream()
        torch.cuda.set_stream(s)
        stream = C.c_void_p(s.cuda_stream)
        L = capi.lib()
        for w in args.workload.split(","):
            fn = {"angular": bench_angular_dist, "mixed": bench_mixed_dist}.get(w)
            if fn is None:
                if dist.get_rank() == 0:
                    print(json.dumps({"workload": w, "skipped": "single-GPU workload (banded: sequential window chain, replicas only)"}), flush=True)
                continue
            line = fn(args, L, stream)
            if line is not None:
                print(json.dumps(line), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return
    torch.cuda.set_device(0)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    stream = C.c_void_p(s.cuda_stream)
    L = capi.lib()
    for w in args.workload.split(","):
        line = {"angular": bench_angular, "mixed": bench_mixed, "two_call": bench_two_call, "classes": bench_classes, "banded": bench_banded, "angular_wide": bench_angular_wide, "lm": bench_lm}[w](args, L, stream)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
