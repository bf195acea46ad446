"""World-size-2 test of the multi-GPU host logic on CPU (gloo): shards partition
the global index range and the Counter all-reduce sums the per-rank histograms.
The per-rank histograms are produced by the oracle here (no GPU); on the GPU box
the same code path runs with NCCL tensors."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    from oracle import oracle
    from tests.beamlines import lens_beamline, lens_table
    from trajectories import _engine as eng
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution

    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert eng.dist_info() == (rank, world)
    lo, hi = eng.shard_range(total, rank, world)
    bl = lens_beamline(lens_table())
    src = oracle.make_source(CeNTREXVelocityDistribution(), CeNTREXPositionDistribution())
    local = oracle.run(bl.elements, src, seed=4, first=lo, n=hi - lo)["counters"]
    t = torch.from_numpy(local.copy())
    eng.allreduce_counts(t)
    np.save(Path(out_dir) / f"rank{rank}.npy", np.stack([local, t.numpy()]))
    dist.destroy_process_group()


def test_shard_and_allreduce_world2(tmp_path):
    import torch.multiprocessing as mp

    from oracle import oracle
    from tests.beamlines import lens_beamline, lens_table
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    total = 200_001
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    np.testing.assert_array_equal(r0[1], r1[1])                 # both ranks hold the merged Counter
    np.testing.assert_array_equal(r0[0] + r1[0], r0[1])
    bl = lens_beamline(lens_table())
    src = oracle.make_source(CeNTREXVelocityDistribution(), CeNTREXPositionDistribution())
    whole = oracle.run(bl.elements, src, seed=4, first=0, n=total)["counters"]
    np.testing.assert_array_equal(r0[1], whole)                 # GPU-count invariance of the result
    assert whole.sum() == total
