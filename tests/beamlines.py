"""Beamlines of the BASELINE.json configs, built with the package's own classes
exactly as the reference's example scripts build them."""
from __future__ import annotations

import numpy as np

from trajectories.centrex import apertures_beamline, lens_beamline, lens_table, spa_beamline  # noqa: F401


def standard_ics(n, seed, sigma_perp=39.5):
    """CeNTREX-shaped initial conditions from a seeded NumPy generator (host side, tests only)."""
    rng = np.random.default_rng(seed)
    ic = np.empty((6, n))
    th, rr = rng.uniform(0, 2 * np.pi, n), np.sqrt(rng.uniform(0, 1, n)) * 0.01
    ic[0], ic[1], ic[2] = rr * np.cos(th), rr * np.sin(th), 0.25 * 0.0254
    ic[3], ic[4], ic[5] = rng.normal(0, sigma_perp, n), rng.normal(0, sigma_perp, n), rng.normal(184, 16, n)
    return ic


def honeycomb_beamline(L=0.05, **mesh_kwargs):
    """A Honeycomb between two apertures (the beamline of tests/golden/honeycomb.npz)."""
    from trajectories.beamline import Beamline
    from trajectories.beamline_elements import CircularAperture, Honeycomb, RectangularAperture

    front = CircularAperture(z0=0.1, L=0.01, d=0.12, name="front")
    mesh = Honeycomb(z0=0.3, L=L, name="Honeycomb", **mesh_kwargs)
    back = RectangularAperture(z0=0.6, L=0.01, w=0.06, h=0.06, name="back")
    return Beamline([front, mesh, back])
