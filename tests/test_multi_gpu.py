"""Two ranks over NCCL: run_simulation shards the global index range and all-reduces the
Counter; the merged result equals the single-GPU run with the same seed (GPU-count invariance)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    for p in (str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    from trajectories.centrex import lens_beamline, lens_table
    from trajectories.trajectory_simulator import TrajectorySimulator

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sim = TrajectorySimulator(device=rank, seed=21, chunk=1 << 20)
    sim.run_simulation(lens_beamline(lens_table()), "r", N_traj=3_000_000, apertures_of_interest=["Detected"], n_jobs=10)
    first_rows = np.array([m.trajectory.x[0] for m in sim.result.molecules])
    keys, vals = list(sim.counter.counter_dict.keys()), list(sim.counter.counter_dict.values())
    # plane probe: same sharding, Counter all-reduced, crossings stay on the owning rank
    xy, v = sim.plane_distributions(lens_beamline(lens_table()), 2.0, elements=["Detected"], N_traj=3_000_000, n_jobs=10)
    assert sim.counter.counter_dict == dict(zip(keys, vals))
    np.savez(Path(out_dir) / f"rank{rank}.npz", keys=np.array(keys), vals=np.array(vals), saved=first_rows, xy=xy, v=v)
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path):
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from trajectories.centrex import lens_beamline, lens_table
    from trajectories.trajectory_simulator import TrajectorySimulator

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    c0, c1 = dict(zip(r0["keys"], r0["vals"])), dict(zip(r1["keys"], r1["vals"]))
    assert c0 == c1 and sum(c0.values()) == 3_000_000                  # every rank holds the merged Counter
    sim = TrajectorySimulator(device=0, seed=21)
    sim.run_simulation(lens_beamline(lens_table()), "r", N_traj=3_000_000, apertures_of_interest=["Detected"], n_jobs=10)
    assert {k: int(v) for k, v in c0.items()} == sim.counter.counter_dict
    # saved molecules stay on the owning rank; together, in rank order, they are the single-GPU list
    both = np.concatenate([r0["saved"].reshape(-1, 3), r1["saved"].reshape(-1, 3)])
    single = np.array([m.trajectory.x[0] for m in sim.result.molecules])
    np.testing.assert_array_equal(both, single)
    xy, v = sim.plane_distributions(lens_beamline(lens_table()), 2.0, elements=["Detected"], N_traj=3_000_000, n_jobs=10)
    np.testing.assert_array_equal(np.concatenate([r0["xy"], r1["xy"]]), xy)
    np.testing.assert_array_equal(np.concatenate([r0["v"], r1["v"]]), v)
    assert xy.shape[0] == single.shape[0] > 100
