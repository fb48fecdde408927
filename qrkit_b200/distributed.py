"""Host-side plumbing of the multi-GPU paths (one process per GPU, torch.distributed).

Block-diagonal problems shard by contiguous ranges of independent diagonal blocks with NO data-path collective
(SURVEY §8e).  Block-angular problems need one exchange: the per-GPU m2 x (m2+1) TSQR triangles are all-gathered
(a few KB: latency only) and every rank merges them redundantly (qrk_angular_merge)."""
from __future__ import annotations

import numpy as np


def block_range(num_blocks: int, world: int, rank: int, align: int = 2):
    """Contiguous range [lo, hi) of diagonal blocks owned by `rank`: balanced to within `align` blocks, every
    boundary a multiple of `align` (keeps 16-byte alignment of each rank's slice of odd-sized-block arrays)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    units = (num_blocks + align - 1) // align
    lo = (units * rank // world) * align
    hi = (units * (rank + 1) // world) * align
    return min(lo, num_blocks), min(hi, num_blocks)


def byte_balanced_ranges(rows, cols, world: int):
    """Mixed block sizes (BASELINE config 5): split by cumulative bytes 8*r*c, not by count.
    Returns a list of `world` contiguous [lo, hi) ranges covering all blocks."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(rows * cols)])
    total = cum[-1]
    bounds = [0]
    for g in range(1, world):
        bounds.append(int(np.searchsorted(cum, total * g / world, side="left")))
    bounds.append(len(rows))
    bounds = np.maximum.accumulate(bounds)
    return [(int(bounds[g]), int(bounds[g + 1])) for g in range(world)]


def all_gather_triangles(tri, group=None):
    """All-gather the per-rank TSQR triangle (1-D float64, identical length on every rank) in rank order.
    Works with any torch.distributed backend (nccl on the GPU box: pass a CUDA tensor; gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    t = tri if isinstance(tri, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(tri, dtype=np.float64))
    world = dist.get_world_size(group)
    out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


def angular_compute_solve(solver, mat_local, b_local, group=None):
    """Distributed BlockAngularSparseQR compute+solve: local factor + border apply + TSQR leaf on this rank's
    blocks, all-gather of the triangles, redundant root merge.  `solver` must have been created with
    world = dist.get_world_size().  Returns this rank's [x1_local ; x2]."""
    solver.compute_solve(mat_local, b_local)
    gathered = all_gather_triangles(solver.local_triangle(), group)
    return solver.merge(gathered.cpu().numpy())


def attach_p2p(handle, group=None, device=None):
    """Fused peer exchange for a block-angular handle (created with qrk_angular_set_world(world) already called): export this
    rank's exchange buffer as a CUDA IPC handle, all-gather the 64-byte handles with torch.distributed, map the peers'
    buffers and hand them to qrk_angular_p2p_attach.  Afterwards qrk_compute_solve / qrk_solve on this handle are complete
    calls again: the triangles travel over NVLink inside the TSQR root kernel (no NCCL on the data path).
    Returns the imported pointers (close them with qrk_ipc_close after destroying the handle)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from .capi import check, lib
    L = lib()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    buf, nbytes = C.c_void_p(), C.c_int64()
    check(L.qrk_angular_xchg_buffer(handle, C.byref(buf), C.byref(nbytes)), handle)
    hd = (C.c_ubyte * 64)()
    check(L.qrk_ipc_export(buf, hd))
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    mine = torch.tensor(list(hd), dtype=torch.uint8, device=dev)
    allh = torch.empty(world * 64, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh, mine, group=group)
    allh = allh.cpu().numpy().reshape(world, 64)
    peers = (C.c_void_p * world)()
    imported = []
    for g in range(world):
        if g == rank:
            peers[g] = buf.value
        else:
            hb = (C.c_ubyte * 64)(*[int(v) for v in allh[g]])
            p = C.c_void_p()
            check(L.qrk_ipc_import(hb, C.byref(p)))
            peers[g] = p.value
            imported.append(p)
    check(L.qrk_angular_p2p_attach(handle, peers, world, rank), handle)
    dist.barrier(group)
    return imported
