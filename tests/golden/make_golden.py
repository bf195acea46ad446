#!/usr/bin/env python
"""Generate the golden fixtures by executing the UNMODIFIED reference source.

Run in the build container only (needs /root/reference; never at test time):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so the
fixtures are produced by importing /root/reference/src/trajectories through the
stub packages in oracle/stubs (matplotlib, h5py, hexalattice, centrex_TlF are
absent from the image) and calling, per molecule,

    m = Molecule(); m.init_trajectory(beamline, x0, v0); beamline.propagate_through(m)

(trajectory_simulator.py:62-69).  The lens acceleration table is injected
through `ElectrostaticLens.a_interp` (electrostatic_lens.py:32,174) as a scipy
`interp1d`, built from the build's rigid-rotor Stark model; the table itself is
stored in the fixture so the tests do not depend on that model.

Outputs (tests/golden/*.npz) record numpy/scipy versions used.
"""
from __future__ import annotations

import importlib.util
import json
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(ROOT / "oracle" / "stubs"))

import numpy as np  # noqa: E402
import scipy  # noqa: E402
from scipy.interpolate import interp1d  # noqa: E402

from trajectories.beamline import Beamline  # noqa: E402
from trajectories.beamline_elements.apertures import (  # noqa: E402
    CircularAperture,
    FieldPlates,
    RectangularAperture,
)
from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens  # noqa: E402
from trajectories.distributions import (  # noqa: E402
    CeNTREXPositionDistribution,
    CeNTREXVelocityDistribution,
    Distribution,
    GaussianPositionDistribution,
)
from trajectories.molecule import Molecule  # noqa: E402
from trajectories.trajectory_simulator import TrajectorySimulator  # noqa: E402

assert "/root/reference" in sys.modules["trajectories"].__file__

_spec = importlib.util.spec_from_file_location(
    "_tlf", ROOT / "centrex-molecule-trajectories_b200" / "trajectories" / "_tlf.py"
)
_tlf = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_tlf)

M = 0.0254


def lens_table(J=2, mJ=0, V=27.6e3, d=1.75 * 0.0254, mass=(204.38 + 19.00) * 1.67e-27):
    return _tlf.lens_acceleration_table(d, V, mass, J, mJ)


def lens_beamline(table):
    """examples/lens_simulation_beamline.py:21-72"""
    fourK = CircularAperture(z0=1.7 * M, L=0.25 * M, d=1 * M, name="4K shield")
    fortyK = CircularAperture(z0=fourK.z1 + 1.25 * M, L=0.25 * M, d=1 * M, name="40K shield")
    bb = CircularAperture(z0=fortyK.z1 + 2.5 * M, L=0.75 * M, d=4 * M, name="BB exit")
    lens = ElectrostaticLens(z0=bb.z1 + 33 * M, L=0.6, name="ES lens")
    lens.a_interp = interp1d(*table)
    fp = FieldPlates(z0=2.43, L=3.0, w=0.02, name="Field plates")
    dr = RectangularAperture(z0=fp.z1 + 39.9 * M, L=0.25 * M, name="DR aperture", w=0.018, h=0.03)
    return Beamline([fourK, fortyK, bb, lens, fp, dr])


def apertures_beamline():
    """The lens beamline without the lens (BASELINE.json configs[0])."""
    fourK = CircularAperture(z0=1.7 * M, L=0.25 * M, d=1 * M, name="4K shield")
    fortyK = CircularAperture(z0=fourK.z1 + 1.25 * M, L=0.25 * M, d=1 * M, name="40K shield")
    bb = CircularAperture(z0=fortyK.z1 + 2.5 * M, L=0.75 * M, d=4 * M, name="BB exit")
    fp = FieldPlates(z0=2.43, L=3.0, w=0.02, name="Field plates")
    dr = RectangularAperture(z0=fp.z1 + 39.9 * M, L=0.25 * M, name="DR aperture", w=0.018, h=0.03)
    return Beamline([fourK, fortyK, bb, fp, dr])


def spa_beamline():
    """examples/SPA/SPA_distributions.py:21-84"""
    fourK = CircularAperture(z0=1.7 * M, L=0.25 * M, d=1 * M, name="4K shield")
    fortyK = CircularAperture(z0=fourK.z1 + 1.25 * M, L=0.25 * M, d=1 * M, name="40K shield")
    bb = CircularAperture(z0=fortyK.z1 + 2.5 * M, L=0.75 * M, d=4 * M, name="BB exit")
    rc_in = CircularAperture(z0=17.36 * M, L=0.125 * M, d=8e-3, name="RC entrance")
    rc_out = CircularAperture(z0=(17.36 + 9) * M, L=0.125 * M, d=8e-3, name="RC exit")
    spa_in = CircularAperture(z0=bb.z1 + 20.5 * M, L=0.375 * M, d=1.75 * M, name="SPA entrance")
    spa_out = CircularAperture(z0=spa_in.z1 + 9.625 * M, L=0.375 * M, d=1.75 * M, name="SPA exit")
    dr_in = CircularAperture(z0=(35.37 + 11) * M, L=0.125 * M, d=150e-3, name="DR entrance")
    laser = RectangularAperture(z0=dr_in.z1 + 3.02 * M, L=2e-3, name="laser", w=0.05, h=0.05)
    return Beamline([fourK, fortyK, bb, rc_in, rc_out, spa_in, spa_out, dr_in, laser])


def fate_names(beamline):
    names = []
    for e in beamline.elements:
        if type(e).__name__ == "ElectrostaticLens":
            cand = ["Lens entrance", "Inside lens"]
        else:
            cand = [e.name]
        for c in cand:
            if c not in names:
                names.append(c)
    names.append("Detected")
    return names


def run_reference(beamline, ic, rows_per_fate=0):
    """Per-molecule reference propagation; returns fates, last rows, row counts, sample rows."""
    names = fate_names(beamline)
    n = ic.shape[1]
    fate = np.empty(n, dtype=np.int8)
    n_rows = np.empty(n, dtype=np.int32)
    alive = np.empty(n, dtype=np.bool_)
    fin = np.empty((10, n))
    kept: dict[int, int] = {}
    row_idx, row_data = [], []
    for i in range(n):
        m = Molecule()
        m.init_trajectory(beamline, ic[0:3, i].copy(), ic[3:6, i].copy())
        beamline.propagate_through(m)
        f = names.index(m.aperture_hit)
        fate[i] = f
        alive[i] = m.alive
        tr = m.trajectory
        k = tr.x.shape[0]
        assert tr.v.shape[0] == k and tr.a.shape[0] == k and tr.t.shape[0] == k == tr.n
        n_rows[i] = k
        fin[0:3, i], fin[3:6, i], fin[6:9, i], fin[9, i] = tr.x[-1], tr.v[-1], tr.a[-1], tr.t[-1]
        if kept.get(f, 0) < rows_per_fate:
            kept[f] = kept.get(f, 0) + 1
            row_idx.append(i)
            row_data.append(np.concatenate([tr.x, tr.v, tr.a, tr.t[:, None]], axis=1))
    out = dict(fate=fate, n_rows=n_rows, alive=alive, fin=fin, fate_names=np.array(names))
    if row_idx:
        out["row_idx"] = np.array(row_idx, dtype=np.int64)
        out["row_off"] = np.cumsum([0] + [r.shape[0] for r in row_data]).astype(np.int64)
        out["rows"] = np.concatenate(row_data, axis=0)
    return out


def draw_reference(seed, n, vdist, xdist, positions_first=False):
    np.random.seed(seed)
    if positions_first:
        xs = xdist.draw(n)
        vs = vdist.draw(n)
    else:  # trajectory_simulator.py:57-58 draws velocities first
        vs = vdist.draw(n)
        xs = xdist.draw(n)
    return np.concatenate([np.asarray(xs, dtype=np.float64), np.asarray(vs, dtype=np.float64)])


def edge_ics(table):
    """Hand-built initial conditions on and next to element edges."""
    R1 = 1 * M / 2            # 4K shield radius
    z4k = 1.7 * M             # its entrance plane
    up = np.nextafter
    rows = [
        # x, y, z, vx, vy, vz
        (0.0, 0.0, 0.00635, 0.0, 0.0, 184.0),               # on axis, no transverse velocity
        (0.0, 0.0, 0.00635, 0.0, 0.0, 250.0),
        (0.0, 0.0, 0.00635, 0.0, 0.055, 184.0),             # launched up against gravity
        (R1, 0.0, z4k, 0.0, 0.0, 184.0),                    # exactly on the 4K edge at dt = 0
        (up(R1, 1.0), 0.0, z4k, 0.0, 0.0, 184.0),           # one ulp outside
        (up(R1, 0.0), 0.0, z4k, 0.0, 0.0, 184.0),           # one ulp inside
        (0.0, R1, z4k, 0.0, 0.0, 184.0),
        (0.0, -R1, z4k, 0.0, 0.0, 184.0),
        (0.009, 0.0, 0.00635, 0.0, 0.1, 184.0),             # reaches field plates with vx == 0
        (0.0099, 0.0, 0.00635, 0.0, 0.1, 184.0),
        (0.0, 0.0, 0.00635, 0.5, 0.05, 184.0),              # slow drift into the field plates (+x)
        (0.0, 0.0, 0.00635, -0.5, 0.05, 184.0),             # (-x)
        (0.0, 0.0, 0.00635, 0.2, 0.2, 150.0),
        (0.0, 0.0, 0.00635, -0.2, 0.3, 210.0),
        (0.001, -0.001, 0.00635, 1.5, -1.0, 184.0),         # enters the lens off axis
        (0.002, 0.002, 0.00635, 3.0, 3.0, 184.0),           # hits the lens bore inside
        (0.0, 0.0, 0.00635, 4.2, 0.0, 184.0),               # near the lens entrance edge
        (0.0, 0.0, 0.00635, 0.0, 4.3, 184.0),
        (0.0, 0.0, 0.00635, 0.3, 0.06, 120.0),              # slow molecule, strong focusing
        (0.0, 0.0, 0.00635, 0.3, 0.06, 300.0),              # fast molecule, weak focusing
    ]
    return np.array(rows, dtype=np.float64).T.copy()


class Replay(Distribution):
    """Hands out consecutive slices of a fixed (3,N) array, like a seeded draw would."""

    def __init__(self, data):
        self.data, self.pos = data, 0

    def draw(self, n):
        out = self.data[:, self.pos:self.pos + n]
        self.pos += n
        return out

    def save_to_hdf(self, *a, **k):
        pass


def plane_fixture(table, meta):
    """post_processing.find_radial_pos_dist / find_vel_dist (post_processing.py:20-140) on reference molecules."""
    from trajectories.post_processing import find_radial_pos_dist, find_vel_dist
    from trajectories.trajectory_simulator import Counter, SimulationResult

    xstd = CeNTREXPositionDistribution()
    ic = draw_reference(321, 400, CeNTREXVelocityDistribution(sigmax=3, sigmay=3), xstd, positions_first=True)
    ic = np.concatenate([ic, edge_ics(table)], axis=1)
    bl = lens_beamline(table)
    names = fate_names(bl)
    mols, fate = [], []
    for i in range(ic.shape[1]):
        m = Molecule()
        m.init_trajectory(bl, ic[0:3, i].copy(), ic[3:6, i].copy())
        bl.propagate_through(m)
        mols.append(m)
        fate.append(names.index(m.aperture_hit))
    result = SimulationResult(Counter(), bl, xstd, None, mols)
    lens = bl.elements[3]
    # before the source (index -1 quirk), on the source plane, between apertures, on element planes (rows lie on
    # or within an ulp of them), inside the lens, after it, inside the field plates, past the detection region
    planes = [0.001, 0.00635, 0.03, bl.elements[0].z0, bl.elements[2].z1, 0.5, lens.z0, 1.2, 1.5, lens.z1, 2.0,
              2.43, 4.0, 5.43, 6.5]
    filters = [None, ["Detected"], ["Detected", "Inside lens", "Field plates"]]
    out = dict(ic=ic, table_r=table[0], table_a=table[1], meta=json.dumps(meta), fate=np.array(fate, dtype=np.int8),
               fate_names=np.array(names), planes=np.array(planes),
               filters=np.array(json.dumps(filters)))
    for p, z in enumerate(planes):
        for f, elements in enumerate(filters):
            out[f"xy_{p}_{f}"] = np.asarray(find_radial_pos_dist(result, z, elements), dtype=np.float64)
            out[f"v_{p}_{f}"] = np.asarray(find_vel_dist(result, z, elements), dtype=np.float64)
    np.savez_compressed(HERE / "plane_crossings.npz", **out)
    print("plane crossings:", {float(z): out[f"xy_{p}_0"].shape[0] for p, z in enumerate(planes)}, flush=True)


def honeycomb_beamline(L=0.05):
    """A Honeycomb between two apertures (no example of the reference uses the element; the geometry is its defaults)."""
    from trajectories.beamline_elements.meshes import Honeycomb

    front = CircularAperture(z0=0.1, L=0.01, d=0.12, name="front")
    mesh = Honeycomb(z0=0.3, L=L, name="Honeycomb")
    back = RectangularAperture(z0=0.6, L=0.01, w=0.06, h=0.06, name="back")
    return Beamline([front, mesh, back])


def honeycomb_fixture(meta):
    """Honeycomb.propagate_through (meshes.py:84-117) executed from the reference source.  hexalattice and
    matplotlib are absent: oracle/stubs restates make_grid and RegularPolygon.contains_point (PARITY UNPINNED
    at that third-party boundary; the reference's own logic around them is what this fixture pins)."""
    import warnings

    warnings.simplefilter("ignore", DeprecationWarning)
    bl = honeycomb_beamline()
    mesh = bl.elements[1]
    rng = np.random.default_rng(2024)
    n = 3000
    ic = np.empty((6, n))
    ic[0], ic[1], ic[2] = rng.uniform(-0.032, 0.032, n), rng.uniform(-0.032, 0.032, n), 0.0
    ic[3], ic[4], ic[5] = rng.normal(0, 3.0, n), rng.normal(0, 3.0, n), rng.normal(184, 16, n)
    # molecules aimed at cell 0 (bottom-left corner) that drift towards cell 1 / the row above inside the mesh:
    # `if not idx` (meshes.py:104) re-assigns the cell at z1 only for them
    x0c, y0c = float(mesh.xcoords[0, 0]), float(mesh.ycoords[0, 0])
    pitch = mesh.cell_wall_length * np.sqrt(3)
    k = 400
    vz = rng.normal(184, 5, k)
    tx, ty = x0c + rng.uniform(-0.001, 0.001, k), y0c + rng.uniform(-0.001, 0.001, k)      # position at z0
    drift = rng.choice([0.0, 1.0], k)[None, :] * np.array([[pitch], [0.0]]) + rng.normal(0, 3e-4, (2, k))
    vx, vy = drift[0] / (mesh.L / vz), drift[1] / (mesh.L / vz)
    t0 = mesh.z0 / vz
    extra = np.array([tx - vx * t0, ty - vy * t0 + 0.5 * 9.80665 * t0 ** 2, np.zeros(k), vx, vy, vz])
    # exactly on cell centres, on vertices and edge mid-points of a cell (boundary of the hit test)
    c = 200
    xc, yc = float(mesh.xcoords[c, 0]), float(mesh.ycoords[c, 0])
    rad = (mesh.cell_wall_length * np.sqrt(3) - mesh.cell_wall_thickness / 2) / 2
    th = 2 * np.pi / 6 * np.arange(6) + np.pi / 2
    pts = [(xc, yc)] + [(xc + rad * np.cos(a), yc + rad * np.sin(a)) for a in th] \
        + [(xc + rad * np.sqrt(3) / 2 * np.cos(a + np.pi / 6), yc + rad * np.sqrt(3) / 2 * np.sin(a + np.pi / 6)) for a in th]
    special = np.array([[px, py, mesh.z0, 0.0, 0.0, 184.0] for px, py in pts]).T
    ic = np.concatenate([ic, extra, special], axis=1)
    res = run_reference(bl, ic, rows_per_fate=2)
    np.savez_compressed(HERE / "honeycomb.npz", ic=ic, meta=json.dumps(meta), n_cell0=k,
                        xcoords=np.asarray(mesh.xcoords), ycoords=np.asarray(mesh.ycoords), nx=mesh.nx, ny=mesh.ny,
                        **{f"hc_{key}": v for key, v in res.items()})
    names = list(res["fate_names"])
    print("honeycomb: fates", dict(zip(names, np.bincount(res["fate"], minlength=len(names)))),
          "| aimed at cell 0:", dict(zip(names, np.bincount(res["fate"][n:n + k], minlength=len(names)))), flush=True)



def table_builder_fixture(meta):
    """The reference's OWN a_r(r) table builder (electrostatic_lens.py:174-213) executed as is:
    `lens.lens_acceleration(x)` with a falsy `a_interp`, in a scratch cwd that holds an empty
    `interpolation_functions/` directory (the builder pickles into it, :212-213).  The only substitution is
    `stark_potential` in the reference module's namespace (centrex_TlF is absent): the build's
    rigid-rotor curve, evaluated for the state the reference passes.  Pins linspace extent, nominal-dr gradient,
    unit conversion and the interp1d hand-over, and the pickle cache on a second call."""
    import os
    import pickle
    import tempfile

    import trajectories.beamline_elements.electrostatic_lens as ref_lens
    from centrex_TlF.states import UncoupledBasisState

    calls = []

    def stark(state, Ezs):
        c = state.find_largest_component()
        calls.append((int(c.J), int(c.mJ), len(Ezs)))
        return _tlf.rigid_rotor_stark_joule(int(c.J), int(c.mJ), Ezs)

    points = [  # (J, mJ, V, d)
        (2, 0, 27.6e3, 1.75 * 0.0254),   # the default lens
        (1, 1, 20e3, 1.75 * 0.0254),
        (3, 0, 30e3, 1.75 * 0.0254),
        (0, 0, 24e3, 1.75 * 0.0254),
        (2, 1, 34e3, 1.5 * 0.0254),      # another bore: other grid length and extent
        (3, 2, 27.6e3, 2.0 * 0.0254),
    ]
    out = dict(meta=json.dumps(meta), points=np.array(points, dtype=np.float64))
    saved = ref_lens.stark_potential
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.mkdir(os.path.join(td, "interpolation_functions"))
        os.chdir(td)
        ref_lens.stark_potential = stark
        try:
            for k, (J, mJ, V, d) in enumerate(points):
                state = 1 * UncoupledBasisState(J=J, mJ=mJ, I1=1 / 2, m1=1 / 2, I2=1 / 2, m2=-1 / 2, Omega=0, P=(-1) ** J,
                                                electronic_state="X")
                lens = ElectrostaticLens(z0=1.0, L=0.6, name="ES lens", d=d, V=V, state=state)
                x = np.array([0.3 * d / 2, -0.2 * d / 2, 1.1])
                a = lens.lens_acceleration(x)
                out[f"r_{k}"], out[f"a_{k}"] = np.asarray(lens.a_interp.x), np.asarray(lens.a_interp.y)
                out[f"x_{k}"], out[f"acc_{k}"] = x, np.asarray(a)
                # second lens of the same configuration: must come from the pickle cache (:180-186), not the builder
                n_calls = len(calls)
                twin = ElectrostaticLens(z0=1.0, L=0.6, name="ES lens", d=d, V=V, state=state)
                a2 = twin.lens_acceleration(x)
                assert len(calls) == n_calls and np.array_equal(a, a2) and np.array_equal(twin.a_interp.y, lens.a_interp.y)
            out["cache_files"] = np.array(sorted(os.listdir("interpolation_functions")))
        finally:
            ref_lens.stark_potential = saved
            os.chdir(cwd)
    np.savez_compressed(HERE / "table_builder.npz", **out)
    print("table builder:", [(c, out[f"r_{k}"].shape[0]) for k, c in enumerate(calls)], flush=True)


def run_simulation_fixtures(table, meta):
    """run_simulation(n_jobs=1) of the reference on replayed draws (Counter keys in order of first occurrence,
    saved list, row counts): the lens beamline saving Detected + Inside lens, the lens beamline saving an EARLY
    fate (molecules stopped by the second aperture and at the lens entrance), and the SPA geometry with its
    Gaussian position source."""
    cases = {}
    vbiased, xstd = CeNTREXVelocityDistribution(sigmax=3, sigmay=3), CeNTREXPositionDistribution()
    ic = draw_reference(7, 1300, vbiased, xstd)
    cases[""] = (lens_beamline(table), ic, 1234, ["Detected", "Inside lens"])
    ic = draw_reference(8, 900, CeNTREXVelocityDistribution(sigmax=12, sigmay=12), xstd)
    cases["early_"] = (lens_beamline(table), ic, 850, ["40K shield", "Lens entrance", "no such element"])
    ic = np.concatenate([draw_reference(9, 700, CeNTREXVelocityDistribution(), GaussianPositionDistribution()),
                         draw_reference(10, 500, CeNTREXVelocityDistribution(sigmax=1.5, sigmay=1.5),
                                        GaussianPositionDistribution(sigmax=1e-3, sigmay=1e-3))], axis=1)
    cases["spa_"] = (spa_beamline(), ic, 1200, ["Detected", "RC exit"])
    out = dict(table_r=table[0], table_a=table[1], meta=json.dumps(meta), n_jobs=1)
    for pre, (bl, ic, n_traj, aoi) in cases.items():
        sim = TrajectorySimulator()
        sim.run_simulation(bl, "golden", vdist=Replay(ic[3:6]), xdist=Replay(ic[0:3]), N_traj=n_traj,
                           apertures_of_interest=aoi, n_jobs=1)
        saved = sim.result.molecules
        out.update({
            pre + "ic": ic, pre + "N_traj": n_traj, pre + "aoi": np.array(aoi),
            pre + "counter_keys": np.array(list(sim.counter.counter_dict.keys())),
            pre + "counter_vals": np.array(list(sim.counter.counter_dict.values()), dtype=np.int64),
            pre + "saved_x0": np.array([m.trajectory.x[0] for m in saved]).T.reshape(3, -1),
            pre + "saved_fate": np.array([m.aperture_hit for m in saved]),
            pre + "saved_n_rows": np.array([m.trajectory.x.shape[0] for m in saved], dtype=np.int32),
            pre + "saved_last": np.array([np.concatenate([m.trajectory.x[-1], m.trajectory.v[-1], m.trajectory.a[-1],
                                                          [m.trajectory.t[-1]]]) for m in saved]).reshape(-1, 10),
            pre + "saved_alive": np.array([m.alive for m in saved]),
            pre + "efficiency": sim.counter.calculate_efficiency()})
        print(f"run_simulation[{pre or 'lens'}]: counter {sim.counter.counter_dict}, saved {len(saved)}", flush=True)
    np.savez_compressed(HERE / "run_simulation.npz", **out)


def main():
    t0 = time.time()
    meta = dict(numpy=np.__version__, scipy=scipy.__version__, python=sys.version.split()[0],
                reference="/root/reference (otimgren/centrex-molecule-trajectories)")
    table = lens_table()
    vstd, xstd = CeNTREXVelocityDistribution(), CeNTREXPositionDistribution()

    # --- the reference's own table builder (python make_golden.py tables / runs regenerate single fixtures) ---
    if sys.argv[1:] in ([], ["tables"]):
        table_builder_fixture(meta)
    if sys.argv[1:] == ["tables"]:
        return
    if sys.argv[1:] == ["runs"]:
        run_simulation_fixtures(table, meta)
        return

    # --- post-processing at planes (python make_golden.py planes regenerates only this fixture) ---
    if sys.argv[1:] != ["honeycomb"]:
        plane_fixture(table, meta)
    if sys.argv[1:] == ["planes"]:
        return

    # --- Honeycomb (python make_golden.py honeycomb regenerates only this fixture) ---
    honeycomb_fixture(meta)
    if sys.argv[1:] == ["honeycomb"]:
        return

    # --- standard distributions, lens beamline + apertures-only beamline (configs 1 and 2) ---
    n_std = 4000
    for seed in (0, 1, 2):
        ic = draw_reference(seed, n_std, vstd, xstd)
        res = run_reference(lens_beamline(table), ic, rows_per_fate=1 if seed == 0 else 0)
        ap = run_reference(apertures_beamline(), ic)
        np.savez_compressed(
            HERE / f"std_seed{seed}.npz", ic=ic, table_r=table[0], table_a=table[1],
            meta=json.dumps(meta), **{f"lens_{k}": v for k, v in res.items()},
            **{f"ap_{k}": v for k, v in ap.items()})
        print(f"std seed {seed}: lens fates {np.bincount(res['fate'])}, ap fates {np.bincount(ap['fate'])}"
              f"  [{time.time() - t0:.0f}s]", flush=True)

    # --- lens-biased set: sigma_perp = 3 m/s sends most molecules into the lens ---
    n_lb = 1500
    ic = draw_reference(123, n_lb, CeNTREXVelocityDistribution(sigmax=3, sigmay=3), xstd,
                        positions_first=True)
    res = run_reference(lens_beamline(table), ic, rows_per_fate=3)
    np.savez_compressed(HERE / "lens_biased.npz", ic=ic, table_r=table[0], table_a=table[1],
                        meta=json.dumps(meta), **{f"lens_{k}": v for k, v in res.items()})
    print(f"lens biased: fates {dict(zip(res['fate_names'], np.bincount(res['fate'], minlength=len(res['fate_names']))))}"
          f"  [{time.time() - t0:.0f}s]", flush=True)

    # --- a second state / voltage (J=3, mJ=0 at 30 kV is field-seeking over part of the range) ---
    table2 = lens_table(J=1, mJ=1, V=20e3)
    res = run_reference(lens_beamline(table2), ic[:, :500])
    np.savez_compressed(HERE / "lens_biased_J1m1_20kV.npz", ic=ic[:, :500], table_r=table2[0],
                        table_a=table2[1], meta=json.dumps(meta),
                        **{f"lens_{k}": v for k, v in res.items()})
    print(f"lens biased J=1 mJ=1 20 kV: fates {np.bincount(res['fate'])}  [{time.time() - t0:.0f}s]", flush=True)

    # --- edge cases ---
    ic = edge_ics(table)
    res = run_reference(lens_beamline(table), ic, rows_per_fate=2)
    ap = run_reference(apertures_beamline(), ic, rows_per_fate=2)
    np.savez_compressed(HERE / "edges.npz", ic=ic, table_r=table[0], table_a=table[1],
                        meta=json.dumps(meta), **{f"lens_{k}": v for k, v in res.items()},
                        **{f"ap_{k}": v for k, v in ap.items()})
    print(f"edges: lens fates {res['fate']}, ap fates {ap['fate']}", flush=True)

    # --- SPA beamline, Gaussian position source (config 4) ---
    ic = draw_reference(5, 4000, vstd, GaussianPositionDistribution())
    # the standard source almost never reaches the laser in 4000 draws: add a collimated batch
    ic2 = draw_reference(6, 1000, CeNTREXVelocityDistribution(sigmax=1.5, sigmay=1.5),
                         GaussianPositionDistribution(sigmax=1e-3, sigmay=1e-3))
    ic = np.concatenate([ic, ic2], axis=1)
    res = run_reference(spa_beamline(), ic, rows_per_fate=2)
    np.savez_compressed(HERE / "spa.npz", ic=ic, meta=json.dumps(meta),
                        **{f"spa_{k}": v for k, v in res.items()})
    print(f"spa: fates {dict(zip(res['fate_names'], np.bincount(res['fate'], minlength=len(res['fate_names']))))}"
          f"  [{time.time() - t0:.0f}s]", flush=True)

    # --- run_simulation itself (n_jobs=1) on replayed draws: Counter + saved list semantics ---
    run_simulation_fixtures(table, meta)
    print(f"  [{time.time() - t0:.0f}s]", flush=True)


if __name__ == "__main__":
    main()
