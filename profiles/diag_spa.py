"""cProfile of configs[3] (SPA beamline, 1e9 molecules, detected trajectories saved) through run_simulation."""
import cProfile, pstats, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")]
import torch
from trajectories.centrex import spa_beamline
from trajectories.distributions import GaussianPositionDistribution
from trajectories.trajectory_simulator import TrajectorySimulator
sim = TrajectorySimulator(seed=2026)
bl = spa_beamline()
sim.run_simulation(bl, "warm", N_traj=int(1e7), apertures_of_interest=["Detected"], n_jobs=9, xdist=GaussianPositionDistribution())
for rep in range(2):
    pr = cProfile.Profile()
    torch.cuda.synchronize(); t = time.perf_counter()
    pr.enable()
    sim.run_simulation(bl, "c4", N_traj=int(1e9), apertures_of_interest=["Detected"], n_jobs=9, xdist=GaussianPositionDistribution())
    pr.disable()
    torch.cuda.synchronize()
    print("configs[3]: %.3f s, %d saved" % (time.perf_counter() - t, len(sim.result.molecules)))
    pstats.Stats(pr).sort_stats("tottime").print_stats(16)
t = time.perf_counter()
sim.run_simulation(bl, "c4n", N_traj=int(1e9), n_jobs=9, xdist=GaussianPositionDistribution())
torch.cuda.synchronize()
print("Counter only: %.3f s" % (time.perf_counter() - t))
pr = cProfile.Profile(); torch.cuda.synchronize(); t = time.perf_counter(); pr.enable()
sim.run_simulation(bl, "c4n", N_traj=int(1e9), n_jobs=9, xdist=GaussianPositionDistribution())
pr.disable(); torch.cuda.synchronize()
print("Counter only again: %.3f s" % (time.perf_counter() - t))
pstats.Stats(pr).sort_stats("tottime").print_stats(8)
