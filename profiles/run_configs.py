#!/usr/bin/env python
"""Run the BASELINE.json configurations end to end through the public API on one B200 and print
one JSON object per configuration (wall clock around run_simulation, as a user sees it).

    python profiles/run_configs.py > profiles/r02_configs.jsonl
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from trajectories.centrex import apertures_beamline, lens_beamline, spa_beamline  # noqa: E402
from trajectories.distributions import GaussianPositionDistribution  # noqa: E402
from trajectories.stark_potential import UncoupledBasisState  # noqa: E402
from trajectories.trajectory_simulator import TrajectorySimulator  # noqa: E402


def timed(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out.update(config=label, seconds=dt, molecules_per_s=out["molecules"] / dt)
    print(json.dumps(out), flush=True)


def summary(sim):
    c = sim.counter.counter_dict
    return dict(molecules=sum(c.values()), counter=c, efficiency=sim.counter.calculate_efficiency(),
                saved=len(sim.result.molecules))


def main():
    sim = TrajectorySimulator(seed=2026)
    sim.run_simulation(lens_beamline(), "warm", N_traj=int(1e6), apertures_of_interest=["Detected"], n_jobs=10)

    def c1():
        sim.run_simulation(apertures_beamline(), "c1", N_traj=int(1e5), apertures_of_interest=["Detected"], n_jobs=10)
        return summary(sim)
    timed("configs[0] apertures-only beamline, 1e5 molecules", c1)

    def c2():
        sim.run_simulation(lens_beamline(), "c2", N_traj=int(1e7), apertures_of_interest=["Detected"], n_jobs=10)
        return summary(sim)
    timed("configs[1] lens beamline (examples/lens_simulation_beamline.py), 1e7 molecules, detected trajectories saved", c2)

    def c3():
        bl = lens_beamline()
        lens = bl.find_element("ES lens")
        eff, total = {}, 0
        states = [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1)]
        for J, mJ in states:
            for V in (20e3, 24e3, 27.6e3, 30e3, 34e3):
                lens.state = 1 * UncoupledBasisState(J=J, mJ=mJ, I1=1 / 2, m1=1 / 2, I2=1 / 2, m2=1 / 2, Omega=0,
                                                     P=(-1) ** J, electronic_state="X")
                lens.V, lens.a_interp = V, None
                sim.run_simulation(bl, f"J={J} mJ={mJ} V={V:.0f}", N_traj=int(1e7), apertures_of_interest=["Detected"], n_jobs=10)
                eff[f"J={J},mJ={mJ},V={V / 1e3:g}kV"] = sim.counter.calculate_efficiency()
                total += sum(sim.counter.counter_dict.values())
        return dict(molecules=total, efficiency=eff, runs=len(eff))
    timed("configs[2] 8 states x 5 voltages x 1e7 molecules, detected trajectories saved", c3)

    def c3s():
        states = [1 * UncoupledBasisState(J=J, mJ=mJ, I1=1 / 2, m1=1 / 2, I2=1 / 2, m2=1 / 2, Omega=0, P=(-1) ** J,
                                          electronic_state="X")
                  for J, mJ in [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1)]]
        res = sim.run_sweep(lens_beamline(), states, [21e3, 25e3, 28.6e3, 31e3, 35e3], N_traj=int(1e7), n_jobs=10)
        return dict(molecules=sum(sum(r.counter.counter_dict.values()) for r in res.values()), runs=len(res),
                    efficiency={f"J={k[0]},mJ={k[1]},V={k[2] / 1e3:g}kV": r.counter.calculate_efficiency() for k, r in res.items()})
    timed("configs[2] as ONE run_sweep call, Counters only, 40 Stark tables (full Hamiltonian) built inside the call", c3s)
    timed("configs[2] as ONE run_sweep call again (Stark curves cached)", c3s)

    def c4():
        sim.run_simulation(spa_beamline(), "c4", N_traj=int(1e9), apertures_of_interest=["Detected"], n_jobs=9,
                           xdist=GaussianPositionDistribution())
        out = summary(sim)
        out["rows_per_saved_trajectory"] = int(sim.result.molecules[0].trajectory.x.shape[0])
        return out
    timed("configs[3] SPA beamline (examples/SPA/SPA_distributions.py), 1e9 molecules, detected trajectories saved", c4)

    def c5():
        sim.run_simulation(lens_beamline(), "c5", N_traj=int(1e10), n_jobs=10)
        return summary(sim)
    timed("configs[4] lens beamline, 1e10 molecules on one GPU (Counter only)", c5)


if __name__ == "__main__":
    main()
