"""The HDF5 layout written by SimulationResult.save_to_hdf (reference
trajectory_simulator.py:218-254 and the per-class save_to_hdf methods),
checked through an in-memory h5py stand-in (h5py/libhdf5 are not in the image)."""
import sys

import numpy as np
import pytest

from tests import fake_h5py
from tests.beamlines import lens_beamline, lens_table


@pytest.fixture()
def h5(monkeypatch):
    fake_h5py.STORE.clear()
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    return fake_h5py


def make_result():
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution
    from trajectories.molecule import Molecule
    from trajectories.trajectory_simulator import Counter, SimulationResult

    bl = lens_beamline(lens_table())
    c = Counter()
    c.increment_counter("4K shield", 5)
    c.increment_counter("Detected", 2)
    rows = np.arange(40, dtype=float).reshape(4, 10)
    mols = [Molecule.from_rows(rows, "Detected", True), Molecule.from_rows(rows[:2] + 100, "Detected", True)]
    return SimulationResult(c, bl, CeNTREXPositionDistribution(), CeNTREXVelocityDistribution(), mols)


def test_layout(h5, tmp_path):
    res = make_result()
    path = tmp_path / "out.hdf"
    run = "lens run/2022"                                   # a run name with '/' nests groups (src/main.py:44)
    res.save_to_hdf(path, run)
    f = h5.File(path, "r")
    g = f[run]
    assert set(g.keys()) == {"beamline", "counter", "position_distribution", "trajectories", "velocity_distribution"}
    assert g["counter"].attrs == {"4K shield": 5, "Detected": 2}
    assert set(g["beamline"].keys()) == {"4K shield", "40K shield", "BB exit", "ES lens", "Field plates", "DR aperture"}
    dr = g["beamline/DR aperture"].attrs
    assert dr["class"] == "RectangularAperture" and dr["w"] == 0.018 and dr["h"] == 0.03
    assert {"name", "z0", "L", "x0", "y0", "z1", "x1", "x2", "y1", "y2"} <= set(dr)
    lens = g["beamline/ES lens"].attrs
    assert lens["class"] == "ElectrostaticLens" and "a_interp" not in lens
    assert lens["V"] == 27.6e3 and lens["d"] == 1.75 * 0.0254
    assert lens["x0"] == "0.0" and lens["y0"] == "0.0"       # quirk: falsy values are stored as repr strings
    assert isinstance(lens["state"], str) and "J = 2" in lens["state"]
    pd = g["position_distribution"].attrs
    assert pd["class"] == "CeNTREXPositionDistribution" and pd["d"] == 0.02 and pd["z"] == 0.25 * 0.0254
    vd = g["velocity_distribution"].attrs
    assert vd["class"] == "CeNTREXVelocityDistribution" and vd["vz"] == 184.0 and vd["sigmaz"] == 16.0
    assert set(g["trajectories"].keys()) == {"molecule_0", "molecule_1"}
    m0 = g["trajectories/molecule_0"]
    assert m0.attrs == {"aperture_hit": "Detected", "alive": True}
    assert m0["x"].shape == (4, 3) and m0["v"].shape == (4, 3) and m0["a"].shape == (4, 3) and m0["t"].shape == (4,)
    assert m0["x"].dtype == np.float64 and m0["t"][3] == 39.0
    assert g["trajectories/molecule_1"]["x"].shape == (2, 3)


def test_existing_run_prompts_before_overwrite(h5, tmp_path, monkeypatch):
    res = make_result()
    path = tmp_path / "out.hdf"
    res.save_to_hdf(path, "r")
    monkeypatch.setattr("builtins.input", lambda prompt="": "n")
    res.save_to_hdf(path, "r")                               # declined: nothing changes, no error
    monkeypatch.setattr("builtins.input", lambda prompt="": "y")
    res.counter.increment_counter("Detected", 1)
    res.save_to_hdf(path, "r")
    assert h5.File(path, "r")["r/counter"].attrs["Detected"] == 3


def test_without_h5py_the_builtin_writer_is_used(tmp_path, monkeypatch):
    """No h5py: the same calls write a real HDF5 file through trajectories._minih5 and read it back."""
    from trajectories import _hdf, _minih5, utils

    monkeypatch.setitem(sys.modules, "h5py", None)
    assert _hdf.h5py() is _minih5
    res = make_result()
    path = tmp_path / "x.hdf"
    res.save_to_hdf(path, "lens run/2022")
    assert path.read_bytes()[:8] == b"\x89HDF\r\n\x1a\n"
    back = utils.import_sim_result_from_hdf(path, "lens run/2022")
    assert back.counter.counter_dict == res.counter.counter_dict
    assert [e.name for e in back.beamline.elements] == [e.name for e in res.beamline.elements]   # Beamline sorts by z0
    assert back.xdist == res.xdist and back.vdist == res.vdist
    for a, b in zip(back.molecules, res.molecules):
        np.testing.assert_array_equal(a.trajectory.x, b.trajectory.x)
        np.testing.assert_array_equal(a.trajectory.a, b.trajectory.a)
        assert a.aperture_hit == b.aperture_hit and a.alive == b.alive


def test_round_trip_through_utils(h5, tmp_path):
    """save_to_hdf -> utils.import_sim_result_from_hdf (reference utils.py:149-159)."""
    from trajectories import utils

    res = make_result()
    path = tmp_path / "rt.hdf"
    res.save_to_hdf(path, "run")
    back = utils.import_sim_result_from_hdf(path, "run")
    assert back.counter.counter_dict == res.counter.counter_dict
    assert [type(e).__name__ for e in back.beamline.elements] == [type(e).__name__ for e in res.beamline.elements]
    for a, b in zip(back.beamline.elements, res.beamline.elements):
        assert (a.name, a.z0, a.L, a.z1) == (b.name, b.z0, b.L, b.z1)
    lens = back.beamline.find_element("ES lens")
    assert lens.V == 27.6e3 and lens.x0 == 0.0 and lens.a_interp is None
    assert back.xdist == res.xdist and back.vdist == res.vdist
    assert len(back.molecules) == 2
    for a, b in zip(back.molecules, res.molecules):
        np.testing.assert_array_equal(a.trajectory.x, b.trajectory.x)
        np.testing.assert_array_equal(a.trajectory.t, b.trajectory.t)
        assert a.aperture_hit == b.aperture_hit and a.alive == b.alive and a.trajectory.n == b.trajectory.n


def test_post_processing_at_a_plane():
    """find_radial_pos_dist / find_vel_dist (reference post_processing.py:20-140)."""
    from trajectories.molecule import Molecule
    from trajectories.post_processing import find_radial_pos_dist, find_vel_dist, take_timestep
    from trajectories.trajectory_simulator import Counter, SimulationResult

    g = 9.80665

    def rows(x0, v0, zs):
        out, x, v, t = [], np.array(x0, float), np.array(v0, float), 0.0
        a = np.array([0.0, -g, 0.0])
        out.append(np.concatenate([x, v, a, [t]]))
        for z in zs:
            dt = (z - x[2]) / v[2]
            x, v = take_timestep(x, v, a, dt)
            t += dt
            out.append(np.concatenate([x, v, a, [t]]))
        return np.array(out)

    m1 = Molecule.from_rows(rows([0.001, 0.0, 0.0], [1.0, 2.0, 200.0], [0.5, 1.0, 2.0]), "Detected", True)
    m2 = Molecule.from_rows(rows([0.0, 0.002, 0.0], [0.0, 0.0, 100.0], [0.5]), "4K shield", False)
    res = SimulationResult(Counter(), None, None, None, [m1, m2])
    at = find_radial_pos_dist(res, 0.75)
    assert at.shape == (1, 2)                                      # m2 stopped at z = 0.5
    dt = 0.25 / 200.0                                              # from the stored row at z = 0.5
    x05, v05 = m1.trajectory.x[1], m1.trajectory.v[1]
    np.testing.assert_allclose(at[0], [x05[0] + v05[0] * dt, x05[1] + v05[1] * dt - g * dt ** 2 / 2], rtol=1e-12)
    v = find_vel_dist(res, 0.75)
    np.testing.assert_allclose(v[0], [1.0, v05[1] - g * dt, 200.0], rtol=1e-12)
    both = find_radial_pos_dist(res, 0.5)                          # z is a stored row for both
    assert both.shape == (2, 2) and both[1, 1] == m2.trajectory.x[1, 1]
    assert find_radial_pos_dist(res, 0.5, elements=["Detected"]).shape == (1, 2)
    assert find_vel_dist(res, 5.0).shape == (0,)
