#!/usr/bin/env python
"""Randomised beamlines against the CPU oracle: random element types, sizes, positions (overlaps included), lens tables
and steps, and a source scaled to the geometry, through every launch mode of the library.

    python profiles/fuzz_geometry.py [--cases 200] [--molecules 20000] [--seed 1] > profiles/r02_fuzz_geometry.json

For every case: fates, Counter, work counters and final rows of propagate_ic with final rows (binary64 walk) against
oracle.propagate; the same molecules asking for fates only (FP32 filter on, all forms) against the same fates; a Philox run
from a random device source against the oracle's Philox run, with the filter on and off, and in contracted arithmetic.
One JSON line per failing case (there should be none) and a summary line."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def random_beamline(rng):
    from trajectories.beamline import Beamline
    from trajectories.beamline_elements.apertures import CircularAperture, FieldPlates, RectangularAperture
    from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens, make_interpolator
    from trajectories.beamline_elements.meshes import Honeycomb

    scale = float(10 ** rng.uniform(-3, -1))                 # transverse scale of the apertures, 1 mm ... 10 cm
    n_el = int(rng.integers(1, 8))
    z = float(rng.uniform(0.01, 0.3))
    elems, lenses, meshes = [], 0, 0
    for k in range(n_el):
        L = float(10 ** rng.uniform(-3, -0.3))
        kind = rng.choice(["circ", "rect", "plates", "lens", "mesh"], p=[0.36, 0.22, 0.14, 0.2, 0.08])
        if kind == "lens" and lenses == 2:
            kind = "circ"
        if kind == "mesh" and meshes == 1:
            kind = "rect"
        size = scale * float(rng.uniform(0.3, 3.0))
        x0, y0 = (float(rng.normal(0, 0.2 * scale)), float(rng.normal(0, 0.2 * scale))) if rng.random() < 0.3 else (0.0, 0.0)
        name = f"{kind}{k}"
        if kind == "circ":
            elems.append(CircularAperture(name=name, z0=z, L=L, d=size, x0=x0, y0=y0))
        elif kind == "rect":
            elems.append(RectangularAperture(name=name, z0=z, L=L, w=size, h=size * float(rng.uniform(0.3, 2)), x0=x0, y0=y0))
        elif kind == "plates":
            elems.append(FieldPlates(name=name, z0=z, L=L, w=size, x0=x0))
        elif kind == "mesh":
            meshes += 1
            wall = size / float(rng.uniform(3, 24))
            elems.append(Honeycomb(name=name, z0=z, L=L, width=size, height=size * float(rng.uniform(0.5, 1.5)),
                                   cell_wall_length=wall, cell_wall_thickness=wall * float(rng.uniform(0.02, 0.3))))
        else:
            lenses += 1
            n_pts = int(rng.integers(8, 400))
            r = np.linspace(0.0, 1.01 * size / 2, n_pts) if rng.random() < 0.8 else \
                np.concatenate([[0.0], np.sort(rng.uniform(0, 1.01 * size / 2, n_pts - 2)), [1.01 * size / 2]])
            if np.any(np.diff(r) <= 0):
                r = np.linspace(0.0, 1.01 * size / 2, n_pts)
            a = -float(10 ** rng.uniform(3, 6)) * r ** float(rng.uniform(0.8, 2.0)) * float(rng.choice([1.0, 1.0, -1.0]))
            dz = L / int(rng.integers(1, 700))
            elems.append(ElectrostaticLens(name=name, z0=z, L=L, d=size, dz=dz, a_interp=make_interpolator(r, a)))
        z += L * float(rng.uniform(-0.3, 1.5)) if rng.random() < 0.25 else L + float(10 ** rng.uniform(-3, -0.5))
        z = max(z, 1e-3)
    return Beamline(elems), scale


def random_ics(rng, n, scale, bl):
    vz = rng.normal(184, 16, n)
    if rng.random() < 0.2:
        vz = np.abs(vz) + 1.0
    z_end = max(e.z0 + e.L for e in bl.elements)
    spread = scale / max(z_end, 0.1) * 184 * float(rng.uniform(0.2, 3.0))          # transverse speed that fills the apertures
    ic = np.empty((6, n))
    th, rr = rng.uniform(0, 2 * np.pi, n), np.sqrt(rng.uniform(0, 1, n)) * scale * float(rng.uniform(0.05, 1.0))
    ic[0], ic[1], ic[2] = rr * np.cos(th), rr * np.sin(th), float(rng.uniform(0, 0.009))
    ic[3], ic[4], ic[5] = rng.normal(0, spread, n), rng.normal(0, spread, n), vz
    # hostile molecules among them: vz = 0 or negative, NaN / inf / huge / tiny / subnormal components, -0.0, a start
    # exactly on the first plane
    k = np.arange(n) % 97
    ic[5, k == 1] = 0.0
    ic[5, k == 2] *= -1
    ic[0, k == 3] = np.nan
    ic[4, k == 4] = np.inf
    ic[5, k == 5] = 1e-300
    ic[5, k == 6] = 1e300
    ic[0, k == 7] = 1e200
    ic[2, k == 8] = min(e.z0 for e in bl.elements)
    ic[3, k == 9] = -0.0
    ic[1, k == 10] = 5e-324
    ic[5, k == 11] = np.nan
    ic[0, k == 12], ic[1, k == 12] = 0.0, 0.0          # on the axis: r = 0 in a lens
    ic[3, k == 13], ic[4, k == 13] = 0.0, 0.0
    return ic


def run_case(torch, oracle, eng, nat, rng, n):
    bl, scale = random_beamline(rng)
    ic = random_ics(rng, n, scale, bl)
    want = oracle.propagate(bl.elements, ic)
    prop = eng.Propagator(bl.elements, 0)
    dev = torch.from_numpy(np.ascontiguousarray(ic)).cuda()
    problems = []
    prop.reset()
    res = prop.propagate_ic(dev, want_fate=True, want_final=True)
    torch.cuda.synchronize()
    fate, fin = res.fate.cpu().numpy(), res.final.cpu().numpy()
    if not np.array_equal(fate, want["fate"]):
        problems.append(f"binary64 walk: {(fate != want['fate']).sum()} fates differ")
    if not np.array_equal(res.counters.cpu().numpy(), want["counters"]):
        problems.append("Counter differs")
    work = res.work.cpu().numpy()
    if want["work"][2] == 0 and not np.array_equal(work[:3], want["work"]):
        problems.append(f"work counters differ: {work[:3].tolist()} vs {want['work'].tolist()}")
    same = fate == want["fate"]
    with np.errstate(invalid="ignore", over="ignore"):
        a, b = fin[:, same], want["fin"][:, same]
        ok = np.isfinite(b)
        rel = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-9)
        odd = ~ok & ~((np.isnan(a) & np.isnan(b)) | (a == b))           # non-finite values must be the same kind
    worst = float(rel.max()) if rel.size else 0.0
    if worst > 1e-11:
        problems.append(f"final rows differ by {worst:.3g}")
    if odd.any():
        problems.append(f"{int(odd.sum())} non-finite final-row values differ in kind")
    # full trajectories of the first molecules through the same beamline
    m = min(n, 194)
    rows, offs, _ = prop.trajectories(dev[:, :m].contiguous())
    w2 = oracle.propagate(bl.elements, ic[:, :m], want_rows=True)
    if not np.array_equal(np.diff(offs), w2["n_rows"]):
        problems.append("row counts of the trajectories differ")
    else:
        for k in range(m):
            g_, w_ = rows[offs[k]:offs[k + 1]], w2["rows"][k, : w2["n_rows"][k]]
            with np.errstate(invalid="ignore", over="ignore"):
                fin_ok = np.isfinite(w_)
                r_ = np.abs(g_[fin_ok] - w_[fin_ok]) / np.maximum(np.abs(w_[fin_ok]), 1e-9)
                bad_kind = ~fin_ok & ~((np.isnan(g_) & np.isnan(w_)) | (g_ == w_))
            if (r_.size and r_.max() > 1e-11) or bad_kind.any():
                problems.append(f"trajectory {k} differs ({float(r_.max()) if r_.size else 0:.3g})")
                break
    decided = 0
    # plane crossings on the device (crossing_kernel) against the package's host mirror of the reference's
    # post-processing on the oracle's rows: random planes, element planes among them
    if "rows" in w2:
        from trajectories.molecule import Molecule
        from trajectories.post_processing import state_at_plane

        z_lo, z_hi = min(e.z0 for e in bl.elements), max(e.z0 + e.L for e in bl.elements)
        planes = list(rng.uniform(0.0, 1.2 * z_hi, 5)) + [bl.elements[0].z0, bl.elements[-1].z0 + bl.elements[-1].L, 0.5 * (z_lo + z_hi)]
        out, valid, _ = prop.plane_crossings(dev[:, :m].contiguous(), planes)
        out, valid = out.cpu().numpy(), valid.cpu().numpy()
        for k in range(m):
            if not np.isfinite(ic[:, k]).all() or not np.isfinite(w2["rows"][k, : w2["n_rows"][k]]).all():
                continue                    # the host mirror's comparisons on NaN rows are not what is under test here
            mol = Molecule.from_rows(w2["rows"][k, : w2["n_rows"][k]], "", True)
            for q, zq in enumerate(planes):
                with np.errstate(all="ignore"):
                    st = state_at_plane(mol, float(zq))
                if (st is not None) != bool(valid[q, k]):
                    problems.append(f"plane {zq:.6g}: molecule {k} reaches it on one side only")
                    break
                if st is not None:
                    ref_v = np.concatenate([st[0][:2], st[1]])
                    with np.errstate(all="ignore"):
                        okv = np.isfinite(ref_v)
                        err = np.abs(out[q, :, k][okv] - ref_v[okv]) / np.maximum(np.abs(ref_v[okv]), 1e-9)
                    if err.size and err.max() > 1e-10:
                        problems.append(f"plane {zq:.6g}: molecule {k} differs by {float(err.max()):.3g}")
                        break
            else:
                continue
            break

    flat = eng.flatten(bl.elements)
    for flags in (0, 4, 8):                 # constant-threshold filter in pairs, per-molecule tolerances, one molecule per thread
        nat.lib().cmt_debug_flags(flags)
        p2 = eng.Propagator(flat, 0)
        p2.dev = eng.DeviceBeamline(flat, 0, "exact")        # the flags are read when a handle is created: bypass the cache
        p2.reset()
        r2 = p2.propagate_ic(dev, want_fate=True, want_final=False)
        torch.cuda.synchronize()
        f2 = r2.fate.cpu().numpy()
        if not np.array_equal(f2, want["fate"]):
            problems.append(f"filter (debug flags {flags}): {(f2 != want['fate']).sum()} fates differ")
        if not np.array_equal(r2.counters.cpu().numpy(), want["counters"]):
            problems.append(f"filter (debug flags {flags}): Counter differs")
        if flags == 0:
            decided = int(r2.work[5])
    nat.lib().cmt_debug_flags(0)

    # The device source on the same geometry: Counter of a Philox run against the oracle's own Philox run (identical
    # integer stream; the transforms differ by ulps between CUDA's and glibc's libm, so a molecule within ~1e-13 of an
    # edge may land on the other side), filter on against filter off on the device (must be identical), and the
    # contracted arithmetic against the exact one.
    from trajectories.distributions import (CeNTREXPositionDistribution, CeNTREXVelocityDistribution,
                                            GaussianPositionDistribution)

    z_end = max(e.z0 + e.L for e in bl.elements)
    spread = scale / max(z_end, 0.1) * 184 * float(rng.uniform(0.2, 3.0))
    vd = CeNTREXVelocityDistribution(sigmax=spread, sigmay=spread * float(rng.uniform(0.5, 2)), vz=float(rng.uniform(100, 300)),
                                     sigmaz=float(rng.uniform(1, 25)))
    zs = float(rng.uniform(0, 0.009))
    xd = CeNTREXPositionDistribution(d=scale * float(rng.uniform(0.1, 2)), z=zs) if rng.random() < 0.5 else \
        GaussianPositionDistribution(sigmax=scale * float(rng.uniform(0.05, 1)), sigmay=scale * float(rng.uniform(0.05, 1)), z=zs)
    seed, first = int(rng.integers(0, 2 ** 62)), int(rng.integers(0, 2 ** 40))
    ref = oracle.run(bl.elements, oracle.make_source(vd, xd), seed, first, n)
    src = eng.make_source(vd, xd)
    runs = {}
    for label, flags, math in (("filter", 0, "exact"), ("no filter", 2, "exact"), ("contracted", 0, "contracted")):
        nat.lib().cmt_debug_flags(flags)
        p3 = eng.Propagator(flat, 0, math=math)
        p3.dev = eng.DeviceBeamline(flat, 0, math)
        p3.reset()
        p3.propagate_philox(src, seed, first, n)
        torch.cuda.synchronize()
        runs[label] = (p3.counters.cpu().numpy().copy(), p3.work.cpu().numpy().copy())
    nat.lib().cmt_debug_flags(0)
    if not np.array_equal(runs["filter"][0], runs["no filter"][0]) or not np.array_equal(runs["filter"][1][:3], runs["no filter"][1][:3]):
        problems.append("Philox run: filter on and off differ")
    moved = int(np.abs(runs["filter"][0] - ref["counters"]).sum()) // 2
    if moved > 2:
        problems.append(f"Philox run: {moved} molecules counted differently from the oracle's run")
    # (a defocusing random table amplifies the 1e-13 differences of the relaxed roundings exponentially along the lens,
    # so more than the odd molecule can change sides there: 5 of 20 000 in the worst of 500 cases)
    moved_c = int(np.abs(runs["contracted"][0] - runs["filter"][0]).sum()) // 2
    if moved_c > max(2, n // 1000):
        problems.append(f"contracted arithmetic: {moved_c} fates differ from the exact run")
    desc = [(type(e).__name__, round(e.z0, 5), round(e.L, 5)) for e in bl.elements]
    return problems, dict(elements=desc, scale=scale, fates=np.bincount(want["fate"], minlength=len(want["counters"])).tolist(),
                          rk_steps=int(want["work"][1]), out_of_range=int(want["work"][2]), filter_decided=decided, worst_rel=worst,
                          philox_moved=moved, contracted_moved=moved_c)


def main():
    import torch
    from oracle import oracle
    from trajectories import _engine as eng
    from trajectories import _native as nat

    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--molecules", type=int, default=20000)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    bad = total_rk = total_decided = with_lens = with_mesh = oob_cases = 0
    worst = 0.0
    distinct, through = [], []
    philox_moved = contracted_moved = 0
    for c in range(args.cases):
        problems, info = run_case(torch, oracle, eng, nat, rng, args.molecules)
        total_rk += info["rk_steps"]
        total_decided += info["filter_decided"]
        with_lens += any(t == "ElectrostaticLens" for t, _, _ in info["elements"])
        with_mesh += any(t == "Honeycomb" for t, _, _ in info["elements"])
        oob_cases += info["out_of_range"] > 0
        worst = max(worst, info["worst_rel"])
        philox_moved += info["philox_moved"]
        contracted_moved += info["contracted_moved"]
        distinct.append(sum(1 for f in info["fates"] if f > 0))
        through.append(info["fates"][-1] / args.molecules if info["fates"] else 0.0)
        if problems:
            bad += 1
            print(json.dumps(dict(case=c, problems=problems, **info)), flush=True)
    print(json.dumps(dict(cases=args.cases, molecules_per_case=args.molecules, seed=args.seed, failing_cases=bad,
                          cases_with_a_lens=with_lens, cases_with_a_honeycomb=with_mesh, cases_with_out_of_table_evaluations=oob_cases, rk_steps=total_rk,
                          fates_decided_by_the_filter=total_decided, worst_relative_difference=worst,
                          philox_molecules_counted_differently_from_the_oracle_run=philox_moved,
                          contracted_fates_different_from_exact=contracted_moved,
                          mean_distinct_fates_per_case=float(np.mean(distinct)),
                          share_of_cases_with_three_or_more_fates=float(np.mean(np.array(distinct) >= 3)))))


if __name__ == "__main__":
    main()
