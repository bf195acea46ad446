"""Rebuild simulation results from the HDF5 layout (reference utils.py:15-159).

Everything is reconstructed from the `class` attribute and the constructor
arguments stored next to it, so files written by the reference and by this
package load the same way.  Needs h5py (optional dependency).
"""
from __future__ import annotations

import importlib
import inspect
from pathlib import Path
from typing import List

import numpy as np

from ._hdf import h5py
from .beamline import Beamline
from .molecule import Molecule, Trajectory
from .trajectory_simulator import Counter, SimulationResult

__all__ = [
    "import_element_from_hdf", "import_beamline_from_hdf", "import_trajectories_from_hdf",
    "import_distribution_from_hdf", "import_counter_from_hdf", "import_sim_result_from_hdf",
]


def _plain(value):
    """h5py hands back numpy scalars / bytes; constructors want Python values."""
    if isinstance(value, bytes):
        return value.decode()
    return value.item() if hasattr(value, "item") and getattr(value, "ndim", 1) == 0 else value


def _rebuild(module_name: str, attrs: dict, drop=()):
    attrs = {k: _plain(v) for k, v in attrs.items()}
    cls = getattr(importlib.import_module(module_name), attrs.pop("class"))
    wanted = [p for p in list(inspect.signature(cls.__init__).parameters)[1:] if p in attrs and p not in drop]
    return cls(**{k: attrs[k] for k in wanted})


def import_element_from_hdf(element_name: str, filepath: Path, run_name: str):
    with h5py().File(filepath, "r") as f:
        attrs = dict(f[run_name + "/beamline/" + element_name].attrs.items())
    # the lens stores its state as a repr string and falsy numbers as strings (electrostatic_lens.py:160-166)
    for key in ("x0", "y0"):
        if isinstance(_plain(attrs.get(key)), str):
            attrs[key] = float(_plain(attrs[key]))
    return _rebuild("trajectories.beamline_elements", attrs, drop=("state", "a_interp"))


def import_beamline_from_hdf(filepath: Path, run_name: str) -> Beamline:
    with h5py().File(filepath, "r") as f:
        names = list(f[run_name + "/beamline"].keys())
    return Beamline([import_element_from_hdf(n, filepath, run_name) for n in names])


def _packed_blocks(grp):
    """The blocks of a packed trajectory group, in global molecule order (one per rank of the run that wrote it)."""
    if "rows" in grp:
        return [grp]
    blocks = [grp[k] for k in grp.keys()]
    return sorted(blocks, key=lambda b: int(_plain(b.attrs["first_molecule"])))


def import_trajectories_from_hdf(filepath: Path, run_name: str, group_name: str = "trajectories") -> List[Molecule]:
    """Molecules of a run: the reference's group-per-molecule layout (utils.py:63-105; molecules come back in the
    group's name order, i.e. molecule_0, molecule_1, molecule_10, ... exactly as with h5py), or, when that group is
    absent, the packed layout written by `save_to_hdf(..., packed=True)` (global molecule order)."""
    molecules = []
    with h5py().File(filepath, "r") as f:
        if group_name == "trajectories" and (run_name + "/trajectories") not in f and (run_name + "/trajectories_packed") in f:
            for block in _packed_blocks(f[run_name + "/trajectories_packed"]):
                rows, offsets = block["rows"][()], block["offsets"][()]
                names = [_plain(n) for n in np.atleast_1d(block.attrs["fate_names"])]
                for k, (fate, alive) in enumerate(zip(block["fate"][()], block["alive"][()])):
                    molecules.append(Molecule.from_rows(rows[offsets[k]:offsets[k + 1]], names[fate], bool(alive)))
            return molecules
        grp = f[run_name + "/" + group_name]
        for name in list(grp.keys()):
            g = grp[name]
            tr = Trajectory(n_rows=0)
            tr.x, tr.v, tr.a, tr.t = g["x"][()], g["v"][()], g["a"][()], g["t"][()]
            tr.n = tr.t.shape[0]
            m = Molecule(alive=bool(_plain(g.attrs["alive"])))
            m.trajectory = tr
            m.set_aperture_hit(_plain(g.attrs["aperture_hit"]))
            molecules.append(m)
    return molecules


def import_distribution_from_hdf(dist_name: str, filepath: Path, run_name: str):
    with h5py().File(filepath, "r") as f:
        attrs = dict(f[run_name + "/" + dist_name].attrs.items())
    return _rebuild("trajectories.distributions", attrs)


def import_counter_from_hdf(filepath: Path, run_name: str) -> Counter:
    with h5py().File(filepath, "r") as f:
        attrs = dict(f[run_name + "/counter"].attrs.items())
    counter = Counter()
    counter.counter_dict = {k: _plain(v) for k, v in attrs.items()}
    return counter


def import_sim_result_from_hdf(filepath: Path, run_name: str) -> SimulationResult:
    return SimulationResult(
        import_counter_from_hdf(filepath, run_name),
        import_beamline_from_hdf(filepath, run_name),
        import_distribution_from_hdf("position_distribution", filepath, run_name),
        import_distribution_from_hdf("velocity_distribution", filepath, run_name),
        import_trajectories_from_hdf(filepath, run_name),
    )
