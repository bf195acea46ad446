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


def test_missing_h5py_is_reported(tmp_path, monkeypatch):
    monkeypatch.setitem(sys.modules, "h5py", None)
    with pytest.raises(ImportError, match="h5py"):
        make_result().save_to_hdf(tmp_path / "x.hdf", "r")
