#!/usr/bin/env python
"""bench.py — headline benchmark of the block-diagonal QR hot path (BASELINE.json configs[1]).

Workload (N = 1): 1,000,000 diagonal blocks of 8x4 FP64 (matrix 8M x 4M), one step = ONE fused pass
"factorize every block + Q^T b + back substitution + column permutation" (compute() + solve(b) of the
reference's BlockDiagonalSparseQR) through the C ABI.  N > 1: every rank owns its own contiguous range of
1M blocks of an N-times larger matrix (weak scaling, no data-path collective: the blocks are independent).

One JSON line on stdout (rank 0):
  value      rows/s with A and b resident in HBM (device pointers through qrk_compute_solve)
  e2e        rows/s through the same C-ABI call with HOST (pinned) buffers: H2D of A and b and D2H of x inside
             the timed region
  roofline   algorithmic bytes (16rc+8r+16c = 640 B/block, +4c with pivoting) / measured kernel time vs the
             measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (restatement of the reference's algorithm; Eigen is absent, the reference itself
             cannot be built) timed on this box's host cores on a bounded sample

  configs    (N = 1) the other BASELINE configs under the same clock: c3 block-angular ellipse Jacobian at 1M points,
             c4 block-banded 100k block rows, c5 100k mixed blocks 32x16..128x64, test4 the reference's wide-border test
             sizes; each with ms_per_step, roofline and a bounded cpu_baseline
  sharded    (N > 1) the paths that shard, timed over the N ranks: strong_8x4 (the named 1M blocks split over N, rotating
             buffers to defeat L2), c3 with its ONE exchange step (NCCL all-gather of the per-GPU triangles vs the fused NVLink
             peer exchange inside the TSQR root kernel), strong and weak, and c5 split by bytes
  sustained_ms_per_step   the headline step over a >= 1 s back-to-back burst (clocks settle under the power cap)

`--impl reference` times the CPU restatement alone (rank 0 only) on the SAME 1M blocks and prints the same line shape.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED_A = 0x51524B49
NB, R, CC = 1_000_000, 8, 4
METRIC = "rows/sec, FP64 block-diagonal QR+solve (1M blocks 8x4 per GPU)"
UNIT = "rows/s"
WORKLOAD = "block-diagonal FP64, 1M blocks of 8x4 per GPU: fused batched Householder QR + Q^T b + back substitution"


def algorithmic_bytes_per_block(r, c, piv):
    return 16 * r * c + 8 * r + 16 * c + (4 * c if piv else 0)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(name):
    """dram bytes per launch of the dominant kernel from the committed ncu summary (profiles/), else None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(name)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


_COMPACT_1T = None


def cpu_reference_run(nb_sample, steps, warmup, piv):
    """The reference's own algorithm (explicit Q_i, sparse Q/R assembly, sparse Q^T b, sparse triangular solve —
    BlockDiagonalSparseQR.h:415-547, 258-280) restated in oracle/, single thread like the reference's serial loop."""
    from helpers import uniform_blocks, vector
    from oracle import oracle as orc
    orc.build()
    vals = uniform_blocks(nb_sample, R, CC)
    b = vector(nb_sample * R, seed=SEED_A + 5)
    times = []
    for i in range(warmup + steps):
        res = orc.bd_reference_uniform(nb_sample, R, CC, vals, b, colpiv=bool(piv))
        if i >= warmup:
            times.append(res["seconds"])
    # the compact variant on all host cores, for context (what a tuned CPU port of OUR algorithm would do)
    threads = os.cpu_count() or 1
    orc.bd_compact_uniform(nb_sample, R, CC, vals, b, colpiv=bool(piv), threads=threads)
    tc = min(orc.bd_compact_uniform(nb_sample, R, CC, vals, b, colpiv=bool(piv), threads=threads)["seconds"] for _ in range(3))
    # SURVEY 8d variant B: the same compact algorithm on ONE thread
    global _COMPACT_1T
    _COMPACT_1T = min(orc.bd_compact_uniform(nb_sample, R, CC, vals, b, colpiv=bool(piv), threads=1)["seconds"] for _ in range(2))
    return sum(times) / len(times), tc, threads


def bench_config(nb, piv, world):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "blocks_per_gpu": nb, "block": [R, CC], "pivoting": "colpiv" if piv else "none",
            "l2": "per-step footprint 640 MB (A 256 + b 64 in, packed 256 + tau 32 + x 32 out) exceeds the 126 MB L2; no flush needed",
            "parallelism": f"{world} x contiguous block ranges, no collective"}


def reference_arm(args):
    """The reference's own algorithm on the host cores, on the SAME per-GPU workload (all 1M blocks, every step).  Eigen is not
    in this image, so the reference cannot be compiled: this is its CPU restatement (oracle/), serial like the reference's
    block loop (BlockDiagonalSparseQR.h:432; its OpenMP pragmas are never enabled).  CPU throughput does not depend on N."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nb_sample = args.blocks
    steps = max(1, min(args.steps, 8))      # ~1.8 s per step: bounded so that the whole run ends within a few minutes
    t, tc, threads = cpu_reference_run(nb_sample, steps, min(args.warmup, 1), args.pivoting)
    value = nb_sample * R / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (counter-based U[0.5,5) blocks, U[-1,1) rhs, generated on device)",
        "config": bench_config(nb_sample, args.pivoting, int(os.environ.get("WORLD_SIZE", "1"))),
        "timed_steps": steps,
        "note": "Eigen is not in this image: the reference cannot be compiled; this is the CPU restatement of its algorithm "
                "(oracle/), serial like the reference's block loop; the same generator as the GPU arm, evaluated on the host",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"all {nb_sample} blocks per step, {steps} timed steps (reference-faithful explicit-Q variant)",
                         "compact_all_cores": {"value": nb_sample * R / tc, "cores": threads},
                         "compact_one_core": {"value": nb_sample * R / _COMPACT_1T, "cores": 1}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_configs(args, L):
    """N = 1: the other BASELINE configs through the same C ABI (bench_extra.py holds the workloads), each with its
    roofline and a bounded cpu_baseline (the oracle on this box's host cores)."""
    import argparse as _ap
    import torch
    import bench_extra as bx
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    a = _ap.Namespace(steps=max(4, min(args.steps, 20)), warmup=3, points=1_000_000, mixed_blocks=100_000, class_blocks=0,
                      banded_blocks=100_000, shapes="", no_cpu=args.no_cpu_baseline, cpu_points=500_000, cpu_mixed_blocks=3000,
                      cpu_banded_blocks=25_000)
    out = {}
    for key, fn in (("c3", bx.bench_angular), ("c5", bx.bench_mixed), ("c4", bx.bench_banded), ("test4", bx.bench_angular_wide)):
        try:
            line = fn(a, L, stream)
            torch.cuda.synchronize()
        except Exception as e:                       # one failing side workload must not take the headline line down
            out[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
            continue
        if key == "c5":                              # straddles the ridge: bound = the larger of the two lower bounds
            rf = line["roofline"]
            t_h, t_f = rf["algorithmic_bytes"] / (rf["hbm"]["peak"] * 1e9), rf["flops"] / (rf["fp64"]["peak"] * 1e12)
            rf["bound"] = "hbm" if t_h >= t_f else "fp64"
            rf["frac"] = max(t_h, t_f) / (line["ms_per_step"] * 1e-3)
            rf["traffic"] = ncu_traffic("bd_wy_factor_kernel (config 5, all classes)")
        elif key == "c3":
            line["roofline"]["traffic"] = ncu_traffic("angular kernels K1+K2+K3 (config 3)")
        elif key == "c4":
            line["roofline"]["traffic"] = None
        out[key] = line
    return out


def run_sharded(args, L, piv, rank, world, local_rank):
    """N > 1: the paths that shard, every rank takes part.  strong_8x4: the named 1M blocks split into contiguous
    ranges, four rotating handle / buffer sets per rank so that a step never finds its data in the 126 MB L2 (one set is
    80 MB at 8 GPUs).  angular_exchange: config 3 with its ONE exchange step, as the NCCL all-gather of the per-GPU triangles
    and as the fused NVLink peer exchange inside the TSQR root kernel, with the 1M points sharded (strong) and per GPU (weak).
    c5_sharded: the 100k mixed blocks split by bytes."""
    import argparse as _ap
    import torch
    import torch.distributed as dist
    import bench_extra as bx
    from qrkit_b200.capi import QRK_DEVICE, QrkDesc, check
    from qrkit_b200.distributed import block_range
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    # ---- strong scaling of the headline shape
    lo, hi = block_range(NB, world, rank)
    nbl = hi - lo
    sets = []
    for k in range(4):
        dA = torch.empty(nbl * R * CC, dtype=torch.float64, device="cuda")
        db = torch.empty(nbl * R, dtype=torch.float64, device="cuda")
        dx = torch.empty(nbl * CC, dtype=torch.float64, device="cuda")
        check(L.qrk_synth_fill(C.c_void_p(dA.data_ptr()), SEED_A + 11 * k, lo, nbl, R, CC, 0.5, 5.0, stream))
        check(L.qrk_synth_fill(C.c_void_p(db.data_ptr()), SEED_A + 5 + 11 * k, lo, nbl, R, 0, -1.0, 1.0, stream))
        d = QrkDesc()
        d.kind, d.device, d.num_blocks, d.block_rows, d.block_cols, d.pivoting = 0, local_rank, nbl, R, CC, piv
        h = C.c_void_p()
        check(L.qrk_create(C.byref(d), C.byref(h)))
        check(L.qrk_set_stream(h, stream), h)
        sets.append((h, dA, db, dx))
    it = [0]

    def step():
        h, dA, db, dx = sets[it[0] & 3]
        it[0] += 1
        check(L.qrk_compute_solve(h, C.c_void_p(dA.data_ptr()), C.c_void_p(db.data_ptr()), C.c_void_p(dx.data_ptr()), QRK_DEVICE), h)
    steps = max(args.steps, 20)
    ms = bx._dist_time(step, steps, 8)
    for (h, *_t) in sets:
        L.qrk_destroy(h)
    del sets
    peak, _src = measured_peaks()
    out["strong_8x4"] = {"workload": f"the named 1M blocks of 8x4 split over {world} GPUs ({nbl} blocks on rank {rank})", "scaling": "strong",
                         "ms_per_step": ms, "value": NB * R / (ms * 1e-3), "unit": UNIT, "steps": steps,
                         "roofline_frac": algorithmic_bytes_per_block(R, CC, piv) * NB / (ms * 1e-3) / 1e9 / (peak * world),
                         "l2": "4 rotating handle + buffer sets per rank (4 x 640/N MB): no step re-reads lines left in L2 by the previous one"}
    # ---- config 3 with its exchange step, config 5 by bytes
    a = _ap.Namespace(steps=max(args.steps, 20), warmup=5, points=1_000_000, mixed_blocks=100_000, scaling="strong", graphs=True)
    ex = {}
    for scaling in ("strong", "weak"):
        a.scaling = scaling
        try:
            line = bx.bench_angular_dist(a, L, stream)
        except Exception as e:
            line = {"error": f"{type(e).__name__}: {e}"[:300]}
        if rank == 0:
            if "error" in line:
                ex[scaling] = line
            else:
                f = line.get("fused_peer_exchange") or {}
                ex[scaling] = {"points_total": 1_000_000 * (world if scaling == "weak" else 1), "nccl_us": line["ms_per_step"] * 1e3,
                               "fused_p2p_us": None if not f else f["ms_per_step"] * 1e3,
                               "x2_identical": bool(line["x2_identical_on_all_ranks"] and f.get("x_identical_to_nccl_path", False)),
                               "rows_per_s_nccl": line["value"], "rows_per_s_fused": f.get("value"),
                               # the same steps replayed from a CUDA graph (kernels + collective captured once): at N > 1 the eager
                               # loop is bound by the host issuing the calls of a ~40 us step, not by the devices
                               "nccl_graph_us": None if not line.get("cuda_graph_replay") else line["cuda_graph_replay"]["ms_per_step"] * 1e3,
                               "fused_p2p_graph_us": None if not f or f.get("cuda_graph_replay_ms_per_step") is None else f["cuda_graph_replay_ms_per_step"] * 1e3,
                               "collective": line["collective"], "roofline_frac_nccl": line["roofline"]["frac"],
                               "roofline_frac_fused": None if not f else 248.0 * (f["value"] / 2) / 1e9 / line["roofline"]["peak"]}
    out["angular_exchange"] = ex
    a.scaling = "strong"
    try:
        out["c5_sharded"] = bx.bench_mixed_dist(a, L, stream)
    except Exception as e:
        out["c5_sharded"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pivoting", type=int, default=0, help="0 = HouseholderQR per block (headline), 1 = ColPivHouseholderQR")
    ap.add_argument("--blocks", type=int, default=NB)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="N = 1: skip the c3 / c4 / c5 / test4 measurements")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip strong_8x4 / angular_exchange / c5_sharded")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from qrkit_b200 import capi
    from qrkit_b200.capi import QRK_DEVICE, QRK_HOST, QrkDesc, check

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: qrkit_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = capi.lib()
    # a non-default torch stream: the handle enqueues on it and torch.cuda.Event records on it (events only see
    # torch's current stream; the legacy default stream has handle 0, which qrk_set_stream reads as "own stream")
    bench_stream = torch.cuda.Stream()
    torch.cuda.set_stream(bench_stream)
    assert bench_stream.cuda_stream != 0
    nb = args.blocks
    piv = args.pivoting

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_handle(pivoting):
        d = QrkDesc()
        d.kind, d.device, d.num_blocks, d.block_rows, d.block_cols, d.pivoting = 0, local_rank, nb, R, CC, pivoting
        h = C.c_void_p()
        check(L.qrk_create(C.byref(d), C.byref(h)))
        check(L.qrk_set_stream(h, C.c_void_p(torch.cuda.current_stream().cuda_stream)), h)
        return h

    # ---- synthetic inputs, generated on the device by the shared counter-based generator (rank r owns blocks
    #      [r*nb, (r+1)*nb) of the N-times larger matrix)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    dA = torch.empty(nb * R * CC, dtype=torch.float64, device="cuda")
    db = torch.empty(nb * R, dtype=torch.float64, device="cuda")
    dx = torch.empty(nb * CC, dtype=torch.float64, device="cuda")
    check(L.qrk_synth_fill(C.c_void_p(dA.data_ptr()), SEED_A, rank * nb, nb, R, CC, 0.5, 5.0, stream))
    check(L.qrk_synth_fill(C.c_void_p(db.data_ptr()), SEED_A + 5, rank * nb, nb, R, 0, -1.0, 1.0, stream))
    torch.cuda.synchronize()

    def launches(h):
        v = C.c_int64()
        L.qrk_launch_count(h, C.byref(v))
        return v.value

    def timed_device(pivoting, steps, warmup, sample_clocks=False):
        h = make_handle(pivoting)

        def step():
            check(L.qrk_compute_solve(h, C.c_void_p(dA.data_ptr()), C.c_void_p(db.data_ptr()), C.c_void_p(dx.data_ptr()), QRK_DEVICE), h)
        for _ in range(warmup):
            step()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
            t_wait = time.time()
            while not sampler.rows and time.time() - t_wait < 5.0:    # nvidia-smi is up and sampling
                time.sleep(0.05)
            sampler.rows.clear()
        barrier()
        l0 = launches(h)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        nl = launches(h) - l0
        clocks = None
        sustained = None
        if sampler:
            # the K-step region lasts a few ms, shorter than one nvidia-smi period: keep the same kernel running
            # back to back for ~1.5 s right behind it so that the sampler sees the clocks under this load; the burst is
            # timed too (sustained_ms_per_step: the figure under the power cap, next to the K-step `value`)
            burst = int(min(50000, max(steps, 1500.0 / max(ms / steps, 1e-3))))
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            for _ in range(burst):
                step()
            b1.record()
            torch.cuda.synchronize()
            sustained = {"ms_per_step": b0.elapsed_time(b1) / burst, "steps": burst}
            clocks = sampler.stop()
            clocks["note"] = f"sampled every 100 ms over the timed region plus a {burst}-step burst of the same kernel behind it"
        raw_ms = ms
        if world > 1:
            t = torch.tensor([ms, sustained["ms_per_step"] if sustained else 0.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0].item())
            if sustained:
                sustained["ms_per_step"] = float(t[1].item())
        L.qrk_destroy(h)
        return ms / steps, nl, clocks, sustained, raw_ms

    ms_step, n_launch, clocks, sustained, raw_elapsed_ms = timed_device(piv, args.steps, args.warmup, sample_clocks=True)
    ms_other = timed_device(1 - piv, args.steps, args.warmup)[0]

    # ---- sanity: the timed step really solved the systems (cheap residual check on a window, on the device)
    h = make_handle(piv)
    check(L.qrk_compute_solve(h, C.c_void_p(dA.data_ptr()), C.c_void_p(db.data_ptr()), C.c_void_p(dx.data_ptr()), QRK_DEVICE), h)
    torch.cuda.synchronize()
    w = 4096
    Aw = dA[: w * R * CC].view(w, CC, R).transpose(1, 2)
    resid = torch.einsum("bij,bj->bi", Aw, dx[: w * CC].view(w, CC)) - db[: w * R].view(w, R)
    normal = torch.einsum("bij,bi->bj", Aw, resid)          # A^T (A x - b) = 0 at the least-squares solution
    ls_check = float(normal.abs().max() / (Aw.abs().max() * db[: w * R].abs().max()))
    L.qrk_destroy(h)
    if not ls_check < 1e-10:
        raise SystemExit(f"bench sanity check failed: normal-equation residual {ls_check}")

    # ---- e2e: the same call with HOST buffers (pinned), H2D of A and b and D2H of x inside the timed region
    e2e = None
    if not args.no_e2e:
        # pinned host buffers, NUMA-local to this rank's GPU (qrk_host_alloc binds the thread to the GPU's PCIe root first)
        numa, ncpu = C.c_int32(-1), C.c_int32(0)
        check(L.qrk_bind_host_thread_to_device(local_rank, C.byref(numa), C.byref(ncpu)))

        def host_tensor(n):
            ptr = C.c_void_p()
            check(L.qrk_host_alloc(C.byref(ptr), n * 8, local_rank))
            arr = np.ctypeslib.as_array((C.c_double * n).from_address(ptr.value))
            return torch.from_numpy(arr), ptr
        (hA, pA), (hb, pb), (hx, px) = host_tensor(nb * R * CC), host_tensor(nb * R), host_tensor(nb * CC)
        hA.copy_(dA); hb.copy_(db)
        h = make_handle(piv)

        def step_host():
            check(L.qrk_compute_solve(h, pA, pb, px, QRK_HOST), h)
        e_steps = max(3, min(args.steps, 10))
        for _ in range(3):
            step_host()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e_steps):
            step_host()
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms_mine = max(e0.elapsed_time(e1), wall) / e_steps   # the call is synchronous: wall time bounds it from above
        ms_e2e = ms_mine
        per_rank = [ms_mine]
        numa_nodes = [int(numa.value)]
        if world > 1:
            t = torch.tensor([ms_mine, float(numa.value)], dtype=torch.float64, device="cuda")
            allt = torch.empty(world * 2, dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(allt, t)
            allt = allt.view(world, 2).cpu()
            per_rank = [float(v) for v in allt[:, 0]]
            numa_nodes = [int(v) for v in allt[:, 1]]
            ms_e2e = max(per_rank)
        assert torch.equal(hx[: w * CC].cuda(), dx[: w * CC]), "host and device paths disagree"
        L.qrk_destroy(h)
        # the platform's ceiling for this step: the same bytes as ONE plain pinned H2D copy per buffer, all ranks at once
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            dA.copy_(hA, non_blocking=True); db.copy_(hb, non_blocking=True)
        c1.record()
        barrier()
        ms_copy = c0.elapsed_time(c1) / 3
        copy_rates = [(hA.numel() + hb.numel()) * 8 / (ms_copy * 1e-3) / 1e9]
        if world > 1:
            t = torch.tensor(copy_rates, dtype=torch.float64, device="cuda")
            allt = torch.empty(world, dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(allt, t)
            copy_rates = [float(v) for v in allt.cpu()]
        h2d = int(hA.numel() * 8 + hb.numel() * 8)
        e2e = {"value": world * nb * R / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(hx.numel() * 8), "ms_per_step": ms_e2e, "steps": e_steps,
               "per_rank_ms_per_step": per_rank, "per_rank_h2d_gbs": [h2d / (m * 1e-3) / 1e9 for m in per_rank],
               "host_numa_node_per_rank": numa_nodes, "host_cpus_bound": int(ncpu.value),
               "plain_h2d_copy_gbs_per_rank": copy_rates,
               "h2d_ceiling_note": "plain cudaMemcpyAsync of the same pinned A and b on every rank at the same time: what the host / PCIe fabric gives N concurrent uploads; the e2e step can not beat it",
               "note": "qrk_compute_solve(QRK_HOST): pinned host A, b -> device, fused kernel, x -> host, per step; the library pipelines the three stages in 16 MB chunks over three streams, so the step is bound by the PCIe upload of A and b; host buffers from qrk_host_alloc (pinned, allocated on the GPU-local NUMA node)"}
        del hA, hb, hx
        for p_ in (pA, pb, px):
            L.qrk_host_free(p_)

    # ---- the other configs (N = 1) and the sharded paths (N > 1) under the same clock
    configs = sharded = None
    if world == 1 and not args.no_configs:
        configs = run_configs(args, L)
    if world > 1 and not args.no_sharded:
        sharded = run_sharded(args, L, piv, rank, world, local_rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    bytes_per_launch = algorithmic_bytes_per_block(R, CC, piv) * nb
    achieved = bytes_per_launch / (ms_step * 1e-3) / 1e9
    kname = f"bd_small_factor_kernel<8,4,{'true' if piv else 'false'},true>"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(kname), "kernel": kname, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "note": "one kernel launch per step; duration = CUDA events over the timed region / steps"}

    cpu = None
    if not args.no_cpu_baseline:
        nb_sample = 250_000
        t, tc, threads = cpu_reference_run(nb_sample, 3, 1, piv)
        cpu = {"value": nb_sample * R / t, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{nb_sample} of {nb} blocks, 3 repetitions; reference-faithful variant (explicit Q_i, sparse Q/R assembly, "
                         "sparse Q^T b + triangular solve), serial like the reference's block loop",
               "compact_all_cores": {"value": nb_sample * R / tc, "cores": threads,
                                     "note": "packed reflectors + fused solve with OpenMP over blocks (not what the reference does)"},
               "compact_one_core": {"value": nb_sample * R / _COMPACT_1T, "cores": 1,
                                    "note": "the same compact algorithm on one thread (SURVEY 8d variant B)"}}

    line = {
        "metric": METRIC, "value": world * nb * R / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (counter-based U[0.5,5) blocks, U[-1,1) rhs, generated on device)",
        "config": bench_config(nb, piv, world),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": n_launch, "clocks": clocks,
        "sustained_ms_per_step": sustained["ms_per_step"] if sustained else None,
        "sustained": None if not sustained else {**sustained, "value": world * nb * R / (sustained["ms_per_step"] * 1e-3),
                                                 "roofline_frac": algorithmic_bytes_per_block(R, CC, piv) * nb / (sustained["ms_per_step"] * 1e-3) / 1e9 / peak,
                                                 "note": "the same step back to back for >= 1 s (the clock sampler's window): the figure under the power cap"},
        "raw_elapsed_ms": raw_elapsed_ms, "configs": configs, "sharded": sharded,
        "other_pivoting": {"pivoting": "none" if piv else "colpiv", "ms_per_step": ms_other,
                           "value": world * nb * R / (ms_other * 1e-3),
                           "roofline_frac": algorithmic_bytes_per_block(R, CC, 1 - piv) * nb / (ms_other * 1e-3) / 1e9 / peak},
        "ls_normal_residual": ls_check,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
