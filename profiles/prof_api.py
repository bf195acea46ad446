#!/usr/bin/env python
"""Where does a steady-state TrajectorySimulator.run_simulation(1e7, apertures_of_interest=["Detected"]) spend its
wall time?  cProfile of the 4th call (bench.py's e2e_api leg)."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")]
import torch  # noqa: E402

from trajectories.centrex import lens_beamline, lens_table  # noqa: E402
from trajectories.trajectory_simulator import TrajectorySimulator  # noqa: E402

bl = lens_beamline(lens_table())
sim = TrajectorySimulator(seed=1)
for k in range(4):
    t = time.perf_counter()
    sim.run_simulation(bl, "r", N_traj=10_000_000, apertures_of_interest=["Detected"], n_jobs=10)
    torch.cuda.synchronize()
    print("call", k, "%.2f ms" % (1e3 * (time.perf_counter() - t)), len(sim.result.molecules), flush=True)
pr = cProfile.Profile()
pr.enable()
sim.run_simulation(bl, "r", N_traj=10_000_000, apertures_of_interest=["Detected"], n_jobs=10)
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(28)
for k in range(3):
    t = time.perf_counter()
    sim.run_simulation(bl, "r", N_traj=10_000_000, n_jobs=10)
    torch.cuda.synchronize()
    print("no saving", k, "%.2f ms" % (1e3 * (time.perf_counter() - t)), flush=True)
for k in range(3):
    sim.run_simulation(bl, "r", N_traj=10_000_000, n_jobs=10)
pr = cProfile.Profile()
pr.enable()
for k in range(10):
    sim.run_simulation(bl, "r", N_traj=10_000_000, n_jobs=10)
pr.disable()
print("ten Counter-only calls:")
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
