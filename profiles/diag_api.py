import gc, sys, time
from pathlib import Path
ROOT = Path("/root/repo")
sys.path[:0] = [str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")]
import torch
from trajectories.centrex import lens_beamline, lens_table
from trajectories.trajectory_simulator import TrajectorySimulator
bl = lens_beamline(lens_table())
sim = TrajectorySimulator(seed=1)
def stats():
    try:
        s = torch.cuda.host_memory_stats()
        return {k: s[k] for k in s if ("alloc" in k and ("count" in k or "num" in k)) or k in ("host_alloc_time.total","host_alloc_time.count")}
    except Exception as e:
        return str(e)
gc.callbacks.append(lambda phase, info: print("   gc", phase, info) if phase == "stop" and info.get("generation", 0) >= 1 else None)
for k in range(30):
    t = time.perf_counter()
    sim.run_simulation(bl, "r", N_traj=10_000_000, apertures_of_interest=["Detected"], n_jobs=10)
    dt = 1e3 * (time.perf_counter() - t)
    print("call", k, "%.2f ms" % dt, flush=True)
    if k in (0, 5, 10, 29): print(stats())
