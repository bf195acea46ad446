"""Electrostatic quadrupole lens.

API mirror of the reference's `beamline_elements/electrostatic_lens.py:23-228`
(fields d, dz, V, a_interp, state, mass; fates "Lens entrance"/"Inside lens").
The RK integration runs on the GPU (csrc/cmt_device.cuh: lens_step); this class
only owns the geometry and the radial-acceleration table a_r(r).
"""
from __future__ import annotations

import pickle
from dataclasses import dataclass, field
from os.path import exists
from typing import Any

import numpy as np

from .apertures import BeamlineElement
from ..stark_potential import default_lens_state, stark_potential, state_quantum_numbers

__all__ = ["ElectrostaticLens"]

INTERP_DIR = "./interpolation_functions/"


class LinearTable:
    """Minimal stand-in for `scipy.interpolate.interp1d(r, a)` (linear, bounds
    checked) used when scipy is not installed; exposes .x/.y like interp1d."""

    def __init__(self, x, y):
        order = np.argsort(x)
        self.x = np.asarray(x, dtype=np.float64)[order]
        self.y = np.asarray(y, dtype=np.float64)[order]

    def __call__(self, r):
        r = np.asarray(r, dtype=np.float64)
        if np.any(r < self.x[0]) or np.any(r > self.x[-1]):
            raise ValueError("A value in x_new is outside the interpolation range.")
        return np.interp(r, self.x, self.y)


def make_interpolator(r_values, a_values):
    try:
        from scipy.interpolate import interp1d

        return interp1d(r_values, a_values)
    except ImportError:  # pragma: no cover
        return LinearTable(r_values, a_values)


@dataclass
class ElectrostaticLens(BeamlineElement):
    d: float = 1.75 * 0.0254            # bore diameter / m
    dz: float = 1e-3                    # integration step along z / m
    V: float = 27.6e3                   # electrode voltage / V
    a_interp: Any = None                # a_r(r) interpolator (scipy interp1d or anything with .x/.y)
    state: Any = field(default_factory=default_lens_state)   # |J=2, mJ=0> by default
    mass: float = (204.38 + 19.00) * 1.67e-27                # TlF mass / kg

    def N_steps(self) -> int:
        return 1 + int(np.rint(self.L / self.dz))

    # -- table --------------------------------------------------------------
    def _cache_name(self) -> str:
        J, mJ = state_quantum_numbers(self.state)
        return f"acceleration_interp_d={self.d:.4f}m_V={self.V:.1f}V_J={J}_mJ={mJ}.pkl"

    def ensure_a_interp(self):
        """Build (or load from ./interpolation_functions/) the a_r(r) interpolator
        the way electrostatic_lens.py:174-213 does; a falsy a_interp triggers it."""
        if self.a_interp:
            return self.a_interp
        path = INTERP_DIR + self._cache_name()
        if exists(path):
            with open(path, "rb") as f:
                self.a_interp = pickle.load(f)
            return self.a_interp
        from .._tlf import lens_acceleration_table

        J, mJ = state_quantum_numbers(self.state)
        r_values, a_values = lens_acceleration_table(
            self.d, self.V, self.mass, J, mJ, stark=lambda Ez: stark_potential(self.state, Ez))
        self.a_interp = make_interpolator(r_values, a_values)
        if exists(INTERP_DIR):
            with open(path, "wb+") as f:
                pickle.dump(self.a_interp, f)
        return self.a_interp

    def acceleration_table(self):
        """(r, a_r) arrays handed to the GPU."""
        f = self.ensure_a_interp()
        if not (hasattr(f, "x") and hasattr(f, "y")):
            raise TypeError("a_interp must expose its table as .x and .y (e.g. scipy.interpolate.interp1d)")
        x = np.asarray(f.x, dtype=np.float64)
        y = np.asarray(f.y, dtype=np.float64)
        if x.ndim != 1 or x.shape != y.shape or x.size < 2:
            raise ValueError("a_interp table must be two 1-D arrays of equal length >= 2")
        if getattr(f, "_kind", "linear") != "linear":
            raise ValueError("only linear a_interp tables are supported on the GPU path")
        return x, y

    def propagate_inside_lens(self, molecule) -> None:
        """The RK integration between the lens' entrance and exit planes for one molecule standing on the entrance
        plane (electrostatic_lens.py:79-118): appends one row per step (a = l1 of the step) and marks the molecule
        "Inside lens" when a step ends beyond the bore.  The reference calls it from `propagate_through` between the
        row at z0 and the row at z1; here it is that same device run (the trajectory kernel on the lens alone, resumed
        from the molecule's last row) with those two rows left out.  A molecule that is not on the entrance plane is
        flown there first, and one that stands outside the bore there is stopped as "Lens entrance", as
        `propagate_through` would (the reference's method tests neither)."""
        from .._single import lens_interior

        lens_interior(self, molecule)

    def lens_acceleration(self, x):
        """Inspection helper: a(x) [m/s^2] at one position from the table
        (electrostatic_lens.py:215-228).  The simulation itself evaluates the
        force inside the CUDA integrator, not through this method."""
        from .._engine import G

        f = self.ensure_a_interp()
        x = np.asarray(x, dtype=np.float64)
        r = np.sqrt(np.sum(x[:2] ** 2))
        a = np.zeros(3)
        if r != 0:
            a_r = float(f(r))
            a[0], a[1] = a_r * x[0] / r, a_r * x[1] / r
        a[1] -= G
        return a

    def save_to_hdf(self, filepath, parent_group_path: str) -> None:
        from .._hdf import save_lens

        save_lens(self, filepath, parent_group_path)
