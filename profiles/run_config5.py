#!/usr/bin/env python
"""BASELINE.json configs[4]: the 1e10-molecule lens-beamline run through the public API, sharded over the
ranks of one torchrun launch (one process per GPU, NCCL all-reduce of the Counter).

    torchrun --nnodes=1 --nproc-per-node N profiles/run_config5.py [N_traj]
"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from trajectories.centrex import lens_beamline  # noqa: E402
from trajectories.trajectory_simulator import TrajectorySimulator  # noqa: E402


def main():
    n_traj = int(float(sys.argv[1])) if len(sys.argv) > 1 else int(1e10)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bl = lens_beamline()
    for math in ("exact", "contracted"):
        sim = TrajectorySimulator(device=local, seed=2026, math=math)
        sim.run_simulation(bl, "warm", N_traj=int(1e7) * world, n_jobs=10)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sim.run_simulation(bl, "config5", N_traj=n_traj, n_jobs=10)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        if rank == 0:
            c = sim.counter.counter_dict
            print(json.dumps(dict(config="configs[4] lens beamline, Counter only", math=math, n_gpus=world,
                                  molecules=sum(c.values()), seconds=dt, molecules_per_s=sum(c.values()) / dt,
                                  counter=c)), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
