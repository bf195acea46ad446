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
    # the merged result file: rank 0 writes the metadata, the ranks append their molecules in turn, numbered globally
    sys.modules["h5py"] = None
    run = TrajectorySimulator(device=rank, seed=21, chunk=1 << 20)
    run.run_simulation(lens_beamline(lens_table()), "r", N_traj=3_000_000, apertures_of_interest=["Detected"], n_jobs=10)
    run.result.save_to_hdf(Path(out_dir) / "merged.hdf", "r")
    # a user-defined Distribution is drawn on the host: identically seeded ranks must simulate disjoint molecules
    from trajectories.distributions import CeNTREXVelocityDistribution

    class Narrow(CeNTREXVelocityDistribution):
        pass

    np.random.seed(99)
    custom = TrajectorySimulator(device=rank)
    custom.run_simulation(lens_beamline(lens_table()), "c", vdist=Narrow(sigmax=3, sigmay=3), N_traj=40_000,
                          apertures_of_interest=["Detected"], n_jobs=1)
    np.savez(Path(out_dir) / f"rank{rank}.npz", keys=np.array(keys), vals=np.array(vals), saved=first_rows, xy=xy, v=v,
             offset=run.result.molecule_offset, total=run.result.n_molecules_total,
             ckeys=np.array(list(custom.counter.counter_dict.keys())), cvals=np.array(list(custom.counter.counter_dict.values())),
             csaved=np.array([m.trajectory.x[0] for m in custom.result.molecules]).reshape(-1, 3))
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
    # the gather of the saved counts: global numbers of each rank's first saved molecule, and one merged file
    assert int(r0["offset"]) == 0 and int(r1["offset"]) == r0["saved"].reshape(-1, 3).shape[0]
    assert int(r0["total"]) == int(r1["total"]) == single.shape[0]
    from trajectories import utils

    sys.modules["h5py"] = None
    sim.result.save_to_hdf(tmp_path / "single.hdf", "r")
    merged = utils.import_trajectories_from_hdf(tmp_path / "merged.hdf", "r")
    alone = utils.import_trajectories_from_hdf(tmp_path / "single.hdf", "r")
    assert len(merged) == len(alone) == single.shape[0]
    for a, b in zip(merged, alone):                                # same molecule_<i> names, same contents
        np.testing.assert_array_equal(a.trajectory.x, b.trajectory.x)
        np.testing.assert_array_equal(a.trajectory.v, b.trajectory.v)
    # host-drawn custom distribution, every rank seeded alike: together the ranks are the single-process run
    from trajectories.distributions import CeNTREXVelocityDistribution

    class Narrow(CeNTREXVelocityDistribution):
        pass

    np.random.seed(99)
    one = TrajectorySimulator(device=0)
    one.run_simulation(lens_beamline(lens_table()), "c", vdist=Narrow(sigmax=3, sigmay=3), N_traj=40_000,
                       apertures_of_interest=["Detected"], n_jobs=1)
    assert {k: int(v) for k, v in zip(r0["ckeys"], r0["cvals"])} == one.counter.counter_dict
    np.testing.assert_array_equal(np.concatenate([r0["csaved"], r1["csaved"]]),
                                  np.array([m.trajectory.x[0] for m in one.result.molecules]).reshape(-1, 3))
