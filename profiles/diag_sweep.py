"""cProfile of a first run_sweep call (40 new Stark tables) and of a repeated one."""
import cProfile, pstats, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")]
import torch
from trajectories.centrex import lens_beamline
from trajectories.trajectory_simulator import TrajectorySimulator
sim = TrajectorySimulator(seed=2026)
bl = lens_beamline()
sim.run_simulation(bl, "warm", N_traj=int(1e6), n_jobs=10)
states = [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1)]
for rep, Vs in enumerate(([20.5e3, 24.5e3, 28.1e3, 30.5e3, 34.5e3], [21.5e3, 25.5e3, 29.1e3, 31.5e3, 35.5e3])):
    pr = cProfile.Profile()
    torch.cuda.synchronize()
    t = time.perf_counter()
    pr.enable()
    sim.run_sweep(bl, states, Vs, N_traj=int(1e7), n_jobs=10)
    pr.disable()
    torch.cuda.synchronize()
    print("first call with new voltages: %.3f s" % (time.perf_counter() - t))
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
t = time.perf_counter(); sim.run_sweep(bl, states, Vs, N_traj=int(1e7), n_jobs=10); torch.cuda.synchronize()
print("repeated: %.3f s" % (time.perf_counter() - t))
