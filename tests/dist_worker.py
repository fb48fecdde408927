"""Worker of tests/test_distributed_cpu.py: run under torch.distributed.run with the gloo backend (world_size 2)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qrkit_b200.distributed import all_gather_triangles, block_range, byte_balanced_ranges  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 1. block ranges tile [0, nb) without overlap, and every rank agrees on them
    nb = 1_000_001
    lo, hi = block_range(nb, world, rank)
    mine = torch.tensor([lo, hi], dtype=torch.int64)
    allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allr, mine)
    assert allr[0][0] == 0 and allr[-1][1] == nb
    for g in range(world - 1):
        assert allr[g][1] == allr[g + 1][0] and allr[g][1] % 2 == 0
    # 2. triangles are gathered in rank order, identically on every rank
    tri = np.arange(20, dtype=np.float64) + 100.0 * rank
    got = all_gather_triangles(tri).numpy()
    want = np.concatenate([np.arange(20, dtype=np.float64) + 100.0 * g for g in range(world)])
    assert np.array_equal(got, want)
    # 3. the max-over-ranks timing reduction used by bench.py
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == float(world)
    # 4. byte-balanced split of mixed block sizes
    rows = 32 + 16 * (np.arange(1000) % 7)
    ranges = byte_balanced_ranges(rows, rows // 2, world)
    assert ranges[0][0] == 0 and ranges[-1][1] == 1000 and all(ranges[g][1] == ranges[g + 1][0] for g in range(world - 1))
    w = [int((rows[a:b] * (rows[a:b] // 2)).sum()) for a, b in ranges]
    assert max(w) - min(w) <= 128 * 64 * 2
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
