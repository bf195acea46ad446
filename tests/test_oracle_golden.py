"""The CPU oracle (oracle/cmt_oracle.c) against the golden vectors produced by
executing the unmodified reference (tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

from oracle import oracle
from tests.beamlines import apertures_beamline, honeycomb_beamline, lens_beamline, spa_beamline


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def check_case(g, prefix, beamline):
    res = oracle.propagate(beamline.elements, g["ic"], want_rows=True)
    assert res["fate_names"] == list(g[f"{prefix}_fate_names"])
    np.testing.assert_array_equal(res["fate"], g[f"{prefix}_fate"])
    np.testing.assert_array_equal(res["n_rows"], g[f"{prefix}_n_rows"])
    np.testing.assert_array_equal(bits(res["fin"]), bits(g[f"{prefix}_fin"]))       # bit for bit
    detected = res["fate"] == res["fate_names"].index("Detected")
    np.testing.assert_array_equal(detected, g[f"{prefix}_alive"])
    if f"{prefix}_rows" in g:
        off = g[f"{prefix}_row_off"]
        for k, i in enumerate(g[f"{prefix}_row_idx"]):
            want = g[f"{prefix}_rows"][off[k]:off[k + 1]]
            got = res["rows"][i, : want.shape[0]]
            np.testing.assert_array_equal(bits(got), bits(want))
            assert np.isnan(res["rows"][i, want.shape[0]:]).all()
    counts = np.bincount(g[f"{prefix}_fate"], minlength=len(res["fate_names"]))
    np.testing.assert_array_equal(res["counters"], counts)
    return res


@pytest.mark.parametrize("name", ["std_seed0", "std_seed1", "std_seed2", "lens_biased", "lens_biased_J1m1_20kV", "edges"])
def test_lens_beamline(golden_dir, name):
    g = np.load(golden_dir / f"{name}.npz")
    res = check_case(g, "lens", lens_beamline((g["table_r"], g["table_a"])))
    assert res["work"][2] == 0   # no force evaluation outside the table


@pytest.mark.parametrize("name", ["std_seed0", "std_seed1", "std_seed2", "edges"])
def test_apertures_only(golden_dir, name):
    g = np.load(golden_dir / f"{name}.npz")
    check_case(g, "ap", apertures_beamline())


def test_spa(golden_dir):
    g = np.load(golden_dir / "spa.npz")
    res = check_case(g, "spa", spa_beamline())
    det = res["fate"] == res["fate_names"].index("Detected")
    assert (res["n_rows"][det] == 19).all()       # 1 + 2 x 9 rows, SURVEY.md 3.4


def test_honeycomb(golden_dir):
    """Honeycomb.propagate_through as the reference executes it (geometry packages restated: parity unpinned there)."""
    g = np.load(golden_dir / "honeycomb.npz")
    res = check_case(g, "hc", honeycomb_beamline())
    names = res["fate_names"]
    n_cell0 = int(g["n_cell0"])
    aimed = res["fate"][3000:3000 + n_cell0]
    assert (aimed == names.index("Detected")).mean() > 0.8      # `if not idx`: cell 0 is re-assigned at z1


def test_row_counts(golden_dir):
    g = np.load(golden_dir / "lens_biased.npz")
    names = list(g["lens_fate_names"])
    det = g["lens_fate"] == names.index("Detected")
    assert det.sum() > 10 and (g["lens_n_rows"][det] == 613).all()


def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [
        ([0, 0, 0, 0], [0, 0], "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         "d16cfe09 94fdcceb 5001e420 24126ea1"),
    ]
    for ctr, key, want in kat:
        got = " ".join(f"{v:08x}" for v in oracle.philox4x32_10(ctr, key))
        assert got == want


def test_source_statistics():
    from trajectories.distributions import (CeNTREXPositionDistribution, CeNTREXVelocityDistribution,
                                            GaussianPositionDistribution)

    n = 400000
    src = oracle.make_source(CeNTREXVelocityDistribution(), CeNTREXPositionDistribution())
    ic = oracle.draw(src, seed=1, first=0, n=n)
    assert abs(ic[3].mean()) < 0.3 and abs(ic[3].std() - 39.5) < 0.2
    assert abs(ic[4].std() - 39.5) < 0.2 and abs(ic[5].mean() - 184) < 0.1 and abs(ic[5].std() - 16) < 0.1
    r = np.hypot(ic[0], ic[1])
    assert r.max() <= 0.01 and abs((r ** 2).mean() - 0.01 ** 2 / 2) < 1e-6    # uniform on the disc
    assert (ic[2] == 0.25 * 0.0254).all()
    # index-addressed: any sub-range reproduces the same samples
    np.testing.assert_array_equal(oracle.draw(src, 1, 1000, 50), ic[:, 1000:1050])
    src = oracle.make_source(CeNTREXVelocityDistribution(), GaussianPositionDistribution())
    ic = oracle.draw(src, seed=2, first=0, n=n)
    s = 0.25 * 25.4 / 5 * 3.8e-3
    assert abs(ic[0].std() - s) < 3e-5 and abs(ic[1].std() - s) < 3e-5
    assert abs(np.corrcoef(ic[0], ic[1])[0, 1]) < 0.01 and abs(np.corrcoef(ic[3], ic[5])[0, 1]) < 0.01


def test_oracle_run_matches_draw_plus_propagate(golden_dir):
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution

    g = np.load(golden_dir / "std_seed0.npz")
    bl = lens_beamline((g["table_r"], g["table_a"]))
    src = oracle.make_source(CeNTREXVelocityDistribution(sigmax=5, sigmay=5), CeNTREXPositionDistribution())
    n = 30000
    a = oracle.run(bl.elements, src, seed=9, first=123, n=n)
    b = oracle.propagate(bl.elements, oracle.draw(src, 9, 123, n))
    np.testing.assert_array_equal(a["counters"], b["counters"])
    np.testing.assert_array_equal(a["work"], b["work"])
    assert a["counters"].sum() == n


def test_plane_crossings_host(golden_dir):
    """post_processing.find_radial_pos_dist / find_vel_dist (reference post_processing.py:20-140): the package's
    host implementation on the oracle's rows reproduces what the reference returned, bit for bit — including the
    index -1 quirk for planes before the source and rows that lie exactly on a plane."""
    import json
    from types import SimpleNamespace

    from trajectories.molecule import Molecule
    from trajectories.post_processing import find_radial_pos_dist, find_vel_dist

    g = np.load(golden_dir / "plane_crossings.npz")
    bl = lens_beamline((g["table_r"], g["table_a"]))
    res = oracle.propagate(bl.elements, g["ic"], want_rows=True)
    np.testing.assert_array_equal(res["fate"], g["fate"])
    names = res["fate_names"]
    mols = [Molecule.from_rows(res["rows"][i, : res["n_rows"][i]], names[res["fate"][i]], alive=names[res["fate"][i]] == "Detected")
            for i in range(g["ic"].shape[1])]
    result = SimpleNamespace(molecules=mols)
    filters = json.loads(str(g["filters"]))
    seen_rows = 0
    for p, z in enumerate(g["planes"]):
        for f, elements in enumerate(filters):
            xy, v = find_radial_pos_dist(result, float(z), elements), find_vel_dist(result, float(z), elements)
            want_xy, want_v = g[f"xy_{p}_{f}"], g[f"v_{p}_{f}"]
            assert xy.shape == want_xy.shape and v.shape == want_v.shape, (z, elements)
            np.testing.assert_array_equal(bits(xy), bits(want_xy))
            np.testing.assert_array_equal(bits(v), bits(want_v))
            seen_rows += want_xy.shape[0]
    assert seen_rows > 5000


def test_table_builder_matches_reference_builder(golden_dir, monkeypatch):
    """L4: `_tlf.lens_acceleration_table` (what ElectrostaticLens.ensure_a_interp calls) against the table the
    reference's OWN builder produced (electrostatic_lens.py:194-209 executed by tests/golden/make_golden.py with only
    `stark_potential` substituted): grid length and extent, the nominal-dr gradient, the division by the mass and
    the interp1d hand-over, bit for bit, for six (state, voltage, bore) points."""
    from trajectories import _tlf
    from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens
    from trajectories.stark_potential import UncoupledBasisState

    from trajectories import stark_potential as sp

    # the fixture was generated with the rigid-rotor curve substituted for centrex_TlF's: the builder is what is pinned
    monkeypatch.setattr(sp, "MODEL", "rigid")
    g = np.load(golden_dir / "table_builder.npz")
    mass = (204.38 + 19.00) * 1.67e-27
    assert len(g["points"]) >= 3
    for k, (J, mJ, V, d) in enumerate(g["points"]):
        r, a = _tlf.lens_acceleration_table(float(d), float(V), mass, int(J), int(mJ))
        np.testing.assert_array_equal(bits(r), bits(g[f"r_{k}"]))
        np.testing.assert_array_equal(bits(a), bits(g[f"a_{k}"]))
        # the same through the package's lens class (falsy a_interp -> build), and its inspection helper
        state = 1 * UncoupledBasisState(J=int(J), mJ=int(mJ), I1=1 / 2, m1=1 / 2, I2=1 / 2, m2=-1 / 2, Omega=0,
                                        P=(-1) ** int(J), electronic_state="X")
        lens = ElectrostaticLens(z0=1.0, L=0.6, name="ES lens", d=float(d), V=float(V), state=state)
        x_tab, y_tab = lens.acceleration_table()
        np.testing.assert_array_equal(bits(x_tab), bits(g[f"r_{k}"]))
        np.testing.assert_array_equal(bits(y_tab), bits(g[f"a_{k}"]))
        np.testing.assert_array_equal(bits(lens.lens_acceleration(g[f"x_{k}"])), bits(g[f"acc_{k}"]))
    # the reference names its pickle cache after d, V, J, mJ (:180-182); the package uses the same names
    lens = ElectrostaticLens(z0=1.0, L=0.6, name="ES lens")
    assert lens._cache_name() in g["cache_files"].tolist()


@pytest.mark.parametrize("case", ["", "early_", "spa_"])
def test_run_simulation_fixtures_vs_oracle(golden_dir, case):
    """The reference's own run_simulation(n_jobs=1) on replayed draws (Counter, saved list in order, row counts,
    last rows) reproduced by the oracle: run size 100*int(N_traj/100), molecules in draw order."""
    full = np.load(golden_dir / "run_simulation.npz")
    g = {k[len(case):]: full[k] for k in full.files if k.startswith(case)} if case else full
    bl = spa_beamline() if case == "spa_" else lens_beamline((full["table_r"], full["table_a"]))
    n = 100 * int(int(g["N_traj"]) / 100)
    res = oracle.propagate(bl.elements, g["ic"][:, :n])
    names = res["fate_names"]
    want = dict(zip(g["counter_keys"].tolist(), g["counter_vals"].tolist()))
    got = {names[f]: int(c) for f, c in enumerate(res["counters"]) if c > 0}
    assert got == want and sum(want.values()) == n
    aoi = set(g["aoi"].tolist())
    keep = np.array([names[f] in aoi for f in res["fate"]])
    assert [names[f] for f in res["fate"][keep]] == g["saved_fate"].tolist()
    np.testing.assert_array_equal(res["n_rows"][keep], g["saved_n_rows"])
    np.testing.assert_array_equal(bits(g["ic"][0:3, :n][:, keep]), bits(g["saved_x0"]))
    np.testing.assert_array_equal(bits(res["fin"][:, keep].T), bits(g["saved_last"]))
