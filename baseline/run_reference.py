#!/usr/bin/env python
"""Times the UNMODIFIED reference (baseline/_ref/trajectories, installed by baseline/install_ref.sh) on the host
cores: `TrajectorySimulator.run_simulation(beamline, run_name, N_traj=N, apertures_of_interest=["Detected"],
n_jobs=k)` (trajectory_simulator.py:34-103) on the full CeNTREX lens beamline of
examples/lens_simulation_beamline.py:21-72, wall clock around the call (sampling and the joblib process pool
included, as a user sees it).  Run by bench.py in a subprocess, because the reference's package is also called
`trajectories`:

    PYTHONPATH=baseline/_ref:oracle/stubs python baseline/run_reference.py TABLE.npz N_TRAJ N_JOBS [N_TRAJ N_JOBS ...]

TABLE.npz holds the lens acceleration table (r, a) -- the same arrays the GPU path gets -- which is injected
through `ElectrostaticLens.a_interp` (electrostatic_lens.py:32,174), so centrex_TlF is neither needed nor timed.
Prints one JSON line: a list of {n_jobs, N_traj, molecules, seconds, molecules_per_s, counter}.
"""
import json
import sys
import time

import numpy as np
from scipy.interpolate import interp1d

from trajectories.beamline import Beamline
from trajectories.beamline_elements.apertures import CircularAperture, FieldPlates, RectangularAperture
from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens
from trajectories.trajectory_simulator import TrajectorySimulator

M = 0.0254


def lens_beamline(table):
    fourK = CircularAperture(z0=1.7 * M, L=0.25 * M, d=1 * M, name="4K shield")
    fortyK = CircularAperture(z0=fourK.z1 + 1.25 * M, L=0.25 * M, d=1 * M, name="40K shield")
    bb = CircularAperture(z0=fortyK.z1 + 2.5 * M, L=0.75 * M, d=4 * M, name="BB exit")
    lens = ElectrostaticLens(z0=bb.z1 + 33 * M, L=0.6, name="ES lens")
    lens.a_interp = interp1d(*table)
    fp = FieldPlates(z0=2.43, L=3.0, w=0.02, name="Field plates")
    dr = RectangularAperture(z0=fp.z1 + 39.9 * M, L=0.25 * M, name="DR aperture", w=0.018, h=0.03)
    return Beamline([fourK, fortyK, bb, lens, fp, dr])


def main():
    import trajectories

    assert "baseline/_ref" in trajectories.__file__.replace("\\", "/"), trajectories.__file__
    tab = np.load(sys.argv[1])
    bl = lens_beamline((tab["r"], tab["a"]))
    out = []
    args = sys.argv[2:]
    for n_traj, n_jobs in zip(args[0::2], args[1::2]):
        n_traj, n_jobs = int(float(n_traj)), int(n_jobs)
        sim = TrajectorySimulator()
        np.random.seed(1)
        t0 = time.perf_counter()
        sim.run_simulation(bl, "bench", N_traj=n_traj, apertures_of_interest=["Detected"], n_jobs=n_jobs)
        dt = time.perf_counter() - t0
        done = sum(sim.counter.counter_dict.values())
        out.append({"n_jobs": n_jobs, "N_traj": n_traj, "molecules": done, "seconds": dt, "molecules_per_s": done / dt,
                    "counter": {k: int(v) for k, v in sim.counter.counter_dict.items()}})
    print("REFERENCE_JSON " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
