#!/usr/bin/env python
"""TrajectorySimulator.run_simulation on randomised beamlines and device sources against the CPU oracle: several chunks per
run (chunk = 2^20), random n_jobs (the reference's run-size arithmetic), random apertures_of_interest; the Counter against
the oracle's Philox run and every saved trajectory, in order, against the oracle's rows for the same global molecule.

    python profiles/fuzz_api.py [--cases 40] [--seed 1] > profiles/r02_fuzz_api.json
"""
import argparse
import importlib.util
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

spec = importlib.util.spec_from_file_location("fuzz_geometry", ROOT / "profiles" / "fuzz_geometry.py")
geo = importlib.util.module_from_spec(spec)
spec.loader.exec_module(geo)


def run_case(rng, oracle, TrajectorySimulator, distributions):
    bl, scale = geo.random_beamline(rng)
    z_end = max(e.z0 + e.L for e in bl.elements)
    spread = scale / max(z_end, 0.1) * 184 * float(rng.uniform(0.2, 3.0))
    vd = distributions.CeNTREXVelocityDistribution(sigmax=spread, sigmay=spread * float(rng.uniform(0.5, 2)),
                                                   vz=float(rng.uniform(120, 300)), sigmaz=float(rng.uniform(1, 20)))
    zs = float(rng.uniform(0, 0.009))
    xd = distributions.CeNTREXPositionDistribution(d=scale * float(rng.uniform(0.1, 2)), z=zs) if rng.random() < 0.5 else \
        distributions.GaussianPositionDistribution(sigmax=scale * float(rng.uniform(0.05, 1)), sigmay=scale * float(rng.uniform(0.05, 1)), z=zs)
    seed = int(rng.integers(0, 2 ** 62))
    n_jobs = int(rng.integers(1, 12))
    N_traj = int(rng.integers(2_200_000, 3_600_000))
    total = int(N_traj / (100 * n_jobs)) * 100 * n_jobs
    src = oracle.make_source(vd, xd)
    ref = oracle.run(bl.elements, src, seed, 0, total)
    names = list(ref["fate_names"])
    small = [nm for nm, c in zip(names, ref["counters"]) if 0 < c <= 15000]
    chosen = list(rng.choice(small, size=min(len(small), int(rng.integers(0, 3))), replace=False)) if small else []
    sim = TrajectorySimulator(seed=seed, chunk=1 << 20)
    problems = []
    try:
        sim.run_simulation(bl, "fuzz", vdist=vd, xdist=xd, N_traj=N_traj, apertures_of_interest=chosen, n_jobs=n_jobs)
    except ValueError as e:                         # the reference raises when a force evaluation leaves the table
        if ref["work"][2] == 0:
            problems.append(f"ValueError without out-of-table evaluations in the oracle's run: {e}")
        return problems, dict(total=total, chosen=chosen, out_of_table=int(ref["work"][2]), saved=0)
    if ref["work"][2] > 0:
        problems.append("the oracle's run left the table but run_simulation did not raise")
    got = sim.counter.counter_dict
    moved = sum(abs(int(got.get(nm, 0)) - int(c)) for nm, c in zip(names, ref["counters"])) // 2
    if sum(got.values()) != total:
        problems.append(f"Counter sums to {sum(got.values())}, not {total}")
    if moved > 2:
        problems.append(f"{moved} molecules counted differently from the oracle's run")
    mols = sim.result.molecules
    want_saved = sum(int(c) for nm, c in zip(names, ref["counters"]) if nm in chosen)
    if abs(len(mols) - want_saved) > 2:
        problems.append(f"{len(mols)} saved molecules, the oracle has {want_saved}")
    worst = 0.0
    if chosen and len(mols) == want_saved and moved == 0:
        ic = oracle.draw(src, seed, 0, total)
        fate = oracle.propagate(bl.elements, ic)["fate"]
        idx = np.nonzero(np.isin(fate, [names.index(c) for c in chosen]))[0]
        step = max(1, len(idx) // 400)
        pick = idx[::step]
        rows = oracle.propagate(bl.elements, ic[:, pick], want_rows=True)
        for j, k in enumerate(range(0, len(idx), step)):
            m = mols[k]
            want = rows["rows"][j, : rows["n_rows"][j]]
            tr = m.trajectory
            have = np.concatenate([tr.x, tr.v, tr.a, tr.t[:, None]], axis=1)
            if m.aperture_hit != names[fate[pick[j]]] or have.shape != want.shape:
                problems.append(f"saved molecule {k}: fate or row count differs")
                break
            with np.errstate(all="ignore"):
                err = np.abs(have - want) / np.maximum(np.abs(want), 1e-9)
            worst = max(worst, float(np.nanmax(err)))
        if worst > 1e-9:
            problems.append(f"saved trajectories differ by {worst:.3g}")
    return problems, dict(total=total, n_jobs=n_jobs, chosen=[str(c) for c in chosen], saved=len(mols), moved=moved,
                          out_of_table=int(ref["work"][2]), worst_rel=worst,
                          elements=[(type(e).__name__, round(e.z0, 5), round(e.L, 5)) for e in bl.elements])


def main():
    from oracle import oracle
    from trajectories import distributions
    from trajectories.trajectory_simulator import TrajectorySimulator

    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    bad = saved = molecules = raised = 0
    worst = 0.0
    for c in range(args.cases):
        problems, info = run_case(rng, oracle, TrajectorySimulator, distributions)
        molecules += info["total"]
        saved += info["saved"]
        raised += info["out_of_table"] > 0
        worst = max(worst, info.get("worst_rel", 0.0))
        if problems:
            bad += 1
            print(json.dumps(dict(case=c, problems=problems, **info)), flush=True)
    print(json.dumps(dict(cases=args.cases, seed=args.seed, failing_cases=bad, molecules=molecules, saved_trajectories=saved,
                          runs_that_raised_for_a_table_excursion=raised, worst_relative_difference_of_a_saved_row=worst)))


if __name__ == "__main__":
    main()
