"""N > 1 host logic on CPU: world_size-2 gloo run of the partitioning and triangle-exchange plumbing
(qrkit_b200/distributed.py) that bench.py and the block-angular multi-GPU path use.  The kernels themselves have no
CPU fallback; the numerical side of the exchange is covered on the GPU by
test_block_angular_gpu.py::test_multi_gpu_exchange_on_one_device."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_range_properties():
    from qrkit_b200.distributed import block_range
    for nb in (0, 1, 7, 1000, 1_000_000, 1_000_001):
        for world in (1, 2, 4, 8):
            rs = [block_range(nb, world, g) for g in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == nb
            assert all(rs[g][1] == rs[g + 1][0] for g in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 2 or nb < 2 * world


def test_gloo_world_size_2():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29591", os.path.join(ROOT, "tests", "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "rank 0 ok" in res.stdout and "rank 1 ok" in res.stdout
