"""Host-side logic that needs no GPU: the API mirror of the reference, flattening,
the C-ABI library's exports, and loud failure without a device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from tests.beamlines import apertures_beamline, lens_beamline, lens_table, spa_beamline

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(cuda_lib):
    """include/cmt.h <-> libcmt_b200.so: every declared entry point is exported (no compute calls)."""
    header = (ROOT / "include" / "cmt.h").read_text()
    declared = set(re.findall(r"\b(cmt_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(cuda_lib, name), f"{name} declared in cmt.h but not exported"
    from trajectories import _native

    assert declared == set(_native.EXPORTS)
    assert cuda_lib.cmt_version() == 101
    assert C.sizeof(_native.Element) == 88 and C.sizeof(_native.Source) == 80 and C.sizeof(_native.Outputs) == 80


def test_abi_argument_validation(cuda_lib):
    """Error behaviour that can be checked without a GPU: bad arguments return CMT_EINVAL, never crash."""
    from trajectories import _native as nat

    out = C.c_void_p()
    assert cuda_lib.cmt_beamline_create(None, 1, None, 0, 2, 1, 9.80665, 0, C.byref(out)) == -1
    assert b"elements" in cuda_lib.cmt_last_error()
    assert cuda_lib.cmt_beamline_create(None, 0, None, 0, 0, 0, 9.80665, 0, C.byref(out)) == -1
    assert cuda_lib.cmt_beamline_create(None, 41, None, 0, 2, 1, 9.80665, 0, C.byref(out)) == -1
    assert cuda_lib.cmt_propagate_ic(None, 1, 0, None, 1, None, None, 0, None) == -1
    assert cuda_lib.cmt_trajectories(None, 1, None, 6, 1, None, 0, None, 1, None, None, None, None) == -1
    assert cuda_lib.cmt_workspace_bytes(None, 10) == 0
    assert cuda_lib.cmt_beamline_max_rows(None) == -1
    cuda_lib.cmt_beamline_destroy(None)


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from trajectories import _native
    from trajectories.molecule import Molecule
    from trajectories.trajectory_simulator import TrajectorySimulator

    bl = lens_beamline(lens_table())
    with pytest.raises(_native.NativeError):
        TrajectorySimulator().run_simulation(bl, "r", N_traj=1000)
    m = Molecule()
    m.init_trajectory(bl)
    with pytest.raises(_native.NativeError):
        bl.propagate_through(m)
    with pytest.raises(_native.NativeError):
        bl.elements[0].propagate_through(m)


def test_product_never_imports_oracle():
    """The shipped package must not reference the oracle (test infrastructure)."""
    for path in (ROOT / "centrex-molecule-trajectories_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".h"}:
            text = path.read_text()
            assert "oracle" not in text.lower(), path


def test_beamline_sorts_in_place_and_finds_elements():
    from trajectories.beamline import Beamline
    from trajectories.beamline_elements.apertures import CircularAperture

    a = CircularAperture(name="b", z0=2.0, L=0.1)
    b = CircularAperture(name="a", z0=1.0, L=0.1)
    lst = [a, b]
    bl = Beamline(lst)
    assert lst[0] is b and bl.elements is lst              # caller's list sorted in place
    assert bl.find_element("b") is a and bl.find_element("zz") is None
    assert a.z1 == 2.1 and a.x0 == 0.0 and a.y0 == 0.0 and a.N_steps() == 2


def test_element_fields_match_reference_defaults():
    from trajectories.beamline_elements import BeamlineElement, CircularAperture, ElectrostaticLens, FieldPlates, RectangularAperture

    r = RectangularAperture(name="r", z0=1, L=1, x0=0.001, y0=-0.002, w=0.018, h=0.03)
    assert (r.x1, r.x2, r.y1, r.y2) == (0.001 - 0.009, 0.001 + 0.009, -0.002 - 0.015, -0.002 + 0.015)
    f = FieldPlates(name="f", z0=1, L=3)
    assert (f.x1, f.x2, f.w) == (-0.01, 0.01, 0.02)
    assert CircularAperture(name="c", z0=0, L=1).d == 0.0254
    lens = ElectrostaticLens(name="l", z0=1, L=0.6)
    assert lens.d == 1.75 * 0.0254 and lens.dz == 1e-3 and lens.V == 27.6e3 and lens.a_interp is None
    assert lens.mass == (204.38 + 19.00) * 1.67e-27 and lens.N_steps() == 601
    c = lens.state.find_largest_component()
    assert (c.J, c.mJ) == (2, 0)
    assert issubclass(ElectrostaticLens, BeamlineElement)
    with pytest.raises(TypeError):
        BeamlineElement(name="x", z0=0, L=1)               # abstract, like the reference


def test_flatten_lens_beamline():
    from trajectories import _engine as eng
    from trajectories import _native as nat

    bl = lens_beamline(lens_table())
    flat = eng.flatten(bl.elements)
    assert flat.fate_names == ["4K shield", "40K shield", "BB exit", "Lens entrance", "Inside lens",
                               "Field plates", "DR aperture", "Detected"]
    assert [e.type for e in flat.elements] == [0, 0, 0, 3, 2, 1]
    assert flat.max_rows == 613                      # rows of a detected molecule, SURVEY.md 4
    lens = flat.elements[3]
    assert lens.n_steps == 600 and lens.R == 1.75 * 0.0254 / 2 and lens.table == 0
    r, a = flat.tables[0]
    assert len(r) == 222 and r[0] == 0.0 and abs(r[-1] - 1.01 * lens.R) < 1e-18
    assert flat.save_mask(["Detected", "Inside lens", "nonexistent"]) == (1 << 7) | (1 << 4)
    assert eng.flatten(spa_beamline().elements).max_rows == 19
    assert eng.flatten(apertures_beamline().elements).fate_names[-1] == "Detected"


def test_flatten_rejects_unknown_elements():
    from trajectories import _engine as eng
    from trajectories.beamline_elements.apertures import BeamlineElement

    class Custom(BeamlineElement):
        def N_steps(self):
            return 1

    with pytest.raises(TypeError, match="no CUDA implementation"):
        eng.flatten([Custom(name="c", z0=0, L=1)])


def test_honeycomb_geometry(golden_dir):
    """Honeycomb (meshes.py:26-82): derived attributes, the cell grid of hexalattice.make_grid as the oracle's
    stand-in lays it out (tests/golden/honeycomb.npz stores what the reference's element held), flattening."""
    import numpy as np

    from oracle import oracle
    from trajectories import _engine as eng
    from trajectories import _native as nat
    from trajectories.beamline_elements import Honeycomb

    g = np.load(golden_dir / "honeycomb.npz")
    h = Honeycomb(name="mesh", z0=0.3, L=0.01)
    assert (h.x1, h.x2, h.y1, h.y2, h.N_steps()) == (-0.0254, 0.0254, -0.0254, 0.0254, 2)
    assert (h.nx, h.ny) == (int(g["nx"]), int(g["ny"])) == (19, 22)
    assert h.xcoords.shape == h.ycoords.shape == (19 * 22, 1)
    np.testing.assert_array_equal(h.xcoords.view(np.int64), g["xcoords"].view(np.int64))     # bit for bit
    np.testing.assert_array_equal(h.ycoords.view(np.int64), g["ycoords"].view(np.int64))
    assert np.abs(h.xcoords).min() == 0.0 and np.abs(h.ycoords).min() == 0.0                 # a cell sits on the origin
    flat = eng.flatten([h])
    e = flat.elements[0]
    assert (e.type, e.n_steps, e.reserved) == (nat.HONEYCOMB, 19, 22)
    o = oracle.flatten([h]).elements[0]                                                       # the oracle derives the same numbers itself
    assert (e.R, e.dz, e.x1, e.y1) == (o["R"], o["dz"], o["x1"], o["y1"])
    # centres as the kernels compute them: ((col + row%2/2) * pitch) - mid_x, ((row * sqrt(3)/2) * pitch) - mid_y
    col, row = np.arange(19 * 22) % 19, np.arange(19 * 22) // 19
    np.testing.assert_array_equal((col + 0.5 * (row % 2)) * e.dz - e.x1, h.xcoords[:, 0])
    np.testing.assert_array_equal((row * (np.sqrt(3) / 2)) * e.dz - e.y1, h.ycoords[:, 0])
    # the unit hexagon hard-coded in csrc/cmt_device.cuh and oracle/cmt_oracle.c is NumPy's
    theta = (2 * np.pi / 6) * np.arange(7) + np.pi / 2.0
    want = [float.fromhex(v) for v in ("0x1.1a62633145c07p-54", "-0x1.bb67ae8584ca9p-1", "-0x1.bb67ae8584cacp-1",
                                       "-0x1.a79394c9e8a0ap-53", "0x1.bb67ae8584ca8p-1", "0x1.bb67ae8584caep-1")]
    assert list(np.cos(theta)[:6]) == want
    want = [float.fromhex(v) for v in ("0x1.0000000000000p+0", "0x1.0000000000003p-1", "-0x1.ffffffffffffbp-2",
                                       "-0x1.0000000000000p+0", "-0x1.0000000000004p-1", "0x1.ffffffffffff3p-2")]
    assert list(np.sin(theta)[:6]) == want
    assert float(np.sqrt(3) / 2) == float.fromhex("0x1.bb67ae8584caap-1")
    for src in ("centrex-molecule-trajectories_b200/csrc/cmt_device.cuh", "oracle/cmt_oracle.c"):
        text = (golden_dir.parent.parent / src).read_text()
        assert "0x1.1a62633145c07p-54" in text and "0x1.ffffffffffff3p-2" in text and "0x1.bb67ae8584caap-1" in text


def test_duplicate_names_share_a_fate():
    from trajectories import _engine as eng
    from trajectories.beamline_elements.apertures import CircularAperture

    flat = eng.flatten([CircularAperture(name="a", z0=0.1, L=0.1), CircularAperture(name="a", z0=0.5, L=0.1)])
    assert flat.fate_names == ["a", "Detected"] and [e.fate for e in flat.elements] == [0, 0]


def test_lens_table_quirks():
    """Table construction follows electrostatic_lens.py:194-206 (nominal-dr gradient, 1.01 R extent)."""
    from trajectories import _tlf

    d, V, mass = 1.75 * 0.0254, 27.6e3, (204.38 + 19.00) * 1.67e-27
    r, a = _tlf.lens_acceleration_table(d, V, mass, 2, 0)
    assert len(r) == int(np.round(d / 2 / 1e-4)) == 222
    assert abs((r[1] - r[0]) - 1.01571e-4) < 1e-8          # true spacing, SURVEY.md 3.3
    E = 2 * V / (d / 2) ** 2 * r / 100
    Vs = _tlf.rigid_rotor_stark_joule(2, 0, E)
    np.testing.assert_array_equal(a, -np.gradient(Vs, 1e-4) / mass)
    assert a[0] != 0.0                                      # one-sided difference at r = 0
    assert (a[5:] < 0).all()                                # J=2, mJ=0 is low-field seeking here: restoring force
    # Stark model sanity: zero-field energies B J(J+1); J=0 is high-field seeking; second-order shift of J=0
    assert abs(_tlf.rigid_rotor_energies_hz(2, 0, [0.0])[0] - 6 * _tlf.B_ROT_HZ) < 1e-3
    e0 = _tlf.rigid_rotor_energies_hz(0, 0, [0.0, 100.0])
    pert = -(_tlf.D_TLF_HZ_PER_V_CM * 100.0) ** 2 / (6 * _tlf.B_ROT_HZ)
    assert abs((e0[1] - e0[0]) / pert - 1) < 1e-3


def test_lens_table_injection_and_cache(tmp_path, monkeypatch):
    from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens, make_interpolator

    monkeypatch.chdir(tmp_path)
    lens = ElectrostaticLens(name="l", z0=1, L=0.6)
    r, a = lens.acceleration_table()                        # built lazily, no cache dir -> nothing written
    assert len(r) == 222 and not (tmp_path / "interpolation_functions").exists()
    (tmp_path / "interpolation_functions").mkdir()
    lens2 = ElectrostaticLens(name="l", z0=1, L=0.6)
    lens2.acceleration_table()
    files = list((tmp_path / "interpolation_functions").iterdir())
    assert [f.name for f in files] == ["acceleration_interp_d=0.0444m_V=27600.0V_J=2_mJ=0.pkl"]
    lens3 = ElectrostaticLens(name="l", z0=1, L=0.6)
    np.testing.assert_array_equal(lens3.acceleration_table()[1], a)   # loaded from the pickle cache
    inj = ElectrostaticLens(name="l", z0=1, L=0.6, a_interp=make_interpolator([0.0, 0.03], [0.0, -300.0]))
    np.testing.assert_array_equal(inj.acceleration_table()[0], [0.0, 0.03])
    acc = inj.lens_acceleration(np.array([0.003, 0.004, 1.0]))
    np.testing.assert_allclose(acc, [-50 * 0.6, -50 * 0.8 - 9.80665, 0.0])
    with pytest.raises(ValueError):
        inj.lens_acceleration(np.array([0.05, 0.0, 1.0]))  # beyond the table, like interp1d's bounds_error


def test_counter_semantics():
    from trajectories.trajectory_simulator import Counter

    c = Counter()
    assert c.calculate_efficiency() == 0
    c.increment_counter("4K shield")
    c.increment_counter("4K shield")
    assert c.calculate_efficiency() == 0 and isinstance(c.calculate_efficiency(), int)
    c.increment_counter("Detected")
    d = Counter()
    d.increment_counter("Detected")
    d.increment_counter("Field plates")
    c.merge_counters([d, Counter()])
    assert c.counter_dict == {"4K shield": 2, "Detected": 2, "Field plates": 1}
    assert c.calculate_efficiency() == 2 / 5


def test_molecule_and_trajectory_containers():
    from trajectories.molecule import Molecule, Trajectory, g

    assert g == 9.80665
    bl = lens_beamline(lens_table())
    m = Molecule()
    m.init_trajectory(bl)
    assert m.trajectory.x.shape == (10 + 2 * 5 + 601, 3) and m.trajectory.n == 1      # 621 rows allocated
    np.testing.assert_array_equal(m.x(), [0, 0, 0])
    np.testing.assert_array_equal(m.a(), [0, -g, 0])
    np.testing.assert_allclose(m.x(0.5), [0, -g * 0.125, 100.0])
    m.update_trajectory(0.5)
    assert m.trajectory.n == 2 and m.t() == 0.5
    m.trajectory.drop_nans()
    assert m.trajectory.x.shape == (2, 3) and m.trajectory.t.shape == (2,)
    rows = np.arange(30, dtype=float).reshape(3, 10)
    mol = Molecule.from_rows(rows, "Detected", True)
    assert mol.trajectory.n == 3 and mol.trajectory.v[1, 0] == 13 and mol.trajectory.t[2] == 29
    np.testing.assert_array_equal(mol.x(), rows[2, 0:3])
    tr = Trajectory(n_rows=1)
    tr.update([0, 0, 0], [0, 0, 1], [0, -g, 0], 0.0)
    tr.extend_rows(rows)
    assert tr.n == 4 and tr.t[3] == 29


def test_wrapped_trajectories_materialise_their_views_lazily():
    """Molecule.from_rows keeps the row block and slices x, v, a, t out of it on first access; afterwards
    they are plain attributes, as in the reference (molecule.py:115-131), whatever is done first."""
    import copy
    import pickle

    from trajectories.molecule import Molecule

    rows = np.arange(40, dtype=float).reshape(4, 10)

    def fresh():
        return Molecule.from_rows(rows.copy(), "Detected", True)

    m = fresh()
    assert m == Molecule(alive=True) and m.aperture_hit == "Detected" and m.trajectory.n == 4
    assert m.trajectory.t.shape == (4,) and m.trajectory.a.shape == (4, 3)           # any of the four first
    assert m.trajectory.x.base is not None                                            # views, not copies
    np.testing.assert_array_equal(fresh().v(), rows[3, 3:6])
    np.testing.assert_array_equal(fresh().trajectory.as_rows()[:, :9], rows[:, :9])
    for clone in (pickle.loads(pickle.dumps(fresh())), copy.deepcopy(fresh())):
        np.testing.assert_array_equal(clone.trajectory.x, rows[:, 0:3])
        assert clone.alive and clone.aperture_hit == "Detected"
    m = fresh()
    m.update_trajectory(0.5)                                                          # grows the arrays, appends a row
    assert m.trajectory.n == 5 and m.t() == rows[3, 9] + 0.5
    m = fresh()
    m.trajectory.drop_nans()
    assert m.trajectory.x.shape == (4, 3)
    with pytest.raises(AttributeError):
        fresh().trajectory.nonexistent


def test_distributions_draw_shapes_and_aliases():
    from trajectories import distributions as D

    np.random.seed(0)
    v = D.CeNTREXVelocityDistribution().draw(1000)
    x = D.CeNTREXPositionDistribution().draw(1000)
    gx = D.GaussianPositionDistribution().draw(1000)
    assert v.shape == x.shape == gx.shape == (3, 1000)
    assert (np.hypot(x[0], x[1]) <= 0.01).all() and (x[2] == 0.25 * 0.0254).all()
    assert abs(v[2].mean() - 184) < 2
    assert D.StandardVelocityDistribution is D.CeNTREXVelocityDistribution
    assert D.StandardPositionDistribution is D.CeNTREXPositionDistribution
    from trajectories.trajectory_simulator import TrajectorySimulator

    assert TrajectorySimulator.run_simulation_parallel is TrajectorySimulator.run_simulation


def test_source_record_only_for_builtin_distributions():
    from trajectories import _engine as eng
    from trajectories import distributions as D

    s = eng.make_source(D.CeNTREXVelocityDistribution(sigmax=3), D.GaussianPositionDistribution())
    assert s.pos_kind == 1 and s.vsigma[0] == 3 and s.p0 == 0.25 * 25.4 / 5 * 3.8e-3

    class Mine(D.CeNTREXVelocityDistribution):
        def draw(self, n):
            return np.zeros((3, n))

    assert eng.make_source(Mine(), D.CeNTREXPositionDistribution()) is None     # custom draw() is replayed


def test_shard_range_partitions_the_index_space():
    from trajectories import _engine as eng

    for total in (0, 1, 7, 1000, 10**10 + 3):
        for world in (1, 2, 3, 8):
            parts = [eng.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_saved_molecules_behave_like_the_reference_list():
    """run_simulation returns its saved molecules as a sequence that makes the Molecule objects on first access
    (trajectories/molecule.py: SavedMolecules); everything a list of Molecule objects is used for must work."""
    import copy
    import pickle

    from trajectories.molecule import Molecule, SavedMolecules

    rng = np.random.default_rng(3)
    rows = rng.normal(size=(12, 10))
    names = ["A", "Detected", "B"]
    s = SavedMolecules()
    assert len(s) == 0 and s == [] and list(s) == [] and not s
    s.add_rows(rows, [0, 2, 7, 12], [1, 0, 1], names)
    extra = Molecule.from_rows(rng.normal(size=(3, 10)), "B", False)
    s.append(extra)
    t = SavedMolecules()
    t.add_rows(rows[:5] + 1.0, [0, 5], [2], names)
    s.extend(t)
    assert len(s) == 5 and bool(s)
    assert [m.aperture_hit for m in s] == ["Detected", "A", "Detected", "B", "B"]
    assert [m.alive for m in s] == [True, False, True, False, False]
    assert s[1].trajectory.n == 5 and s[-1].trajectory.n == 5 and s[3] is extra
    np.testing.assert_array_equal(s[1].trajectory.x, rows[2:7, 0:3])
    np.testing.assert_array_equal(s[1].trajectory.t, rows[2:7, 9])
    np.testing.assert_array_equal(s[4].trajectory.v, rows[:5, 3:6] + 1.0)
    assert s[0] is s[0] and s[::2][1] is s[2] and isinstance(s[1:3], list) and len(s[1:3]) == 2
    s[2].alive = False                              # a molecule, once made, is the object the caller keeps seeing
    assert [m.alive for m in s][2] is False
    with pytest.raises(IndexError):
        s[5]
    assert s == list(s) and s != [] and (s + [extra])[5] is extra and len(s) == 5
    for clone in (pickle.loads(pickle.dumps(s)), copy.deepcopy(s)):
        assert len(clone) == 5 and [m.aperture_hit for m in clone] == [m.aperture_hit for m in s]
        np.testing.assert_array_equal(clone[1].trajectory.a, s[1].trajectory.a)

    # list operations that rearrange: the sequence becomes one plain list first
    v = SavedMolecules()
    v.add_rows(rows, [0, 2, 7, 12], [1, 0, 1], names)
    first, second, third = v[0], v[1], v[2]
    assert v.pop() is third and len(v) == 2
    v.insert(0, extra)
    assert v[0] is extra and v[1] is first and v[2] is second and len(v) == 3
    v.sort(key=lambda m: m.trajectory.n)
    assert [m.trajectory.n for m in v] == [2, 3, 5]
    v.reverse()
    del v[0]
    v[0] = third
    assert len(v) == 2 and v[0] is third and v[1] is first and v.index(first) == [third, first].index(first)
    v.clear()
    assert len(v) == 0 and v == []

    # rows poisoned by a non-finite value are stripped when the molecule is made, as Beamline.propagate_through does
    bad = rows.copy()
    bad[5:7] = np.nan
    u = SavedMolecules()
    u.add_rows(bad, [0, 2, 7, 12], [1, 0, 1], names, strip_nans=True)
    assert u[1].trajectory.n == 3 and u[1].trajectory.x.shape == (3, 3) and u[0].trajectory.n == 2
