"""Beamlines of the BASELINE.json configs, built with the package's own classes
exactly as the reference's example scripts build them."""
from __future__ import annotations

import numpy as np

from trajectories import _tlf
from trajectories.beamline import Beamline
from trajectories.beamline_elements.apertures import CircularAperture, FieldPlates, RectangularAperture
from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens, make_interpolator

M = 0.0254


def lens_table(J=2, mJ=0, V=27.6e3, d=1.75 * 0.0254, mass=(204.38 + 19.00) * 1.67e-27):
    return _tlf.lens_acceleration_table(d, V, mass, J, mJ)


def _front():
    fourK = CircularAperture(z0=1.7 * M, L=0.25 * M, d=1 * M, name="4K shield")
    fortyK = CircularAperture(z0=fourK.z1 + 1.25 * M, L=0.25 * M, d=1 * M, name="40K shield")
    bb = CircularAperture(z0=fortyK.z1 + 2.5 * M, L=0.75 * M, d=4 * M, name="BB exit")
    return fourK, fortyK, bb


def lens_beamline(table=None, V=27.6e3):
    """examples/lens_simulation_beamline.py:21-72 (BASELINE.json configs[1])."""
    fourK, fortyK, bb = _front()
    lens = ElectrostaticLens(z0=bb.z1 + 33 * M, L=0.6, name="ES lens", V=V)
    if table is not None:
        lens.a_interp = make_interpolator(*table)
    fp = FieldPlates(z0=2.43, L=3.0, w=0.02, name="Field plates")
    dr = RectangularAperture(z0=fp.z1 + 39.9 * M, L=0.25 * M, name="DR aperture", w=0.018, h=0.03)
    return Beamline([fourK, fortyK, bb, lens, fp, dr])


def apertures_beamline():
    """The same beamline without the lens (BASELINE.json configs[0])."""
    fourK, fortyK, bb = _front()
    fp = FieldPlates(z0=2.43, L=3.0, w=0.02, name="Field plates")
    dr = RectangularAperture(z0=fp.z1 + 39.9 * M, L=0.25 * M, name="DR aperture", w=0.018, h=0.03)
    return Beamline([fourK, fortyK, bb, fp, dr])


def spa_beamline():
    """examples/SPA/SPA_distributions.py:21-84 (BASELINE.json configs[3])."""
    fourK, fortyK, bb = _front()
    rc_in = CircularAperture(z0=17.36 * M, L=0.125 * M, d=8e-3, name="RC entrance")
    rc_out = CircularAperture(z0=(17.36 + 9) * M, L=0.125 * M, d=8e-3, name="RC exit")
    spa_in = CircularAperture(z0=bb.z1 + 20.5 * M, L=0.375 * M, d=1.75 * M, name="SPA entrance")
    spa_out = CircularAperture(z0=spa_in.z1 + 9.625 * M, L=0.375 * M, d=1.75 * M, name="SPA exit")
    dr_in = CircularAperture(z0=(35.37 + 11) * M, L=0.125 * M, d=150e-3, name="DR entrance")
    laser = RectangularAperture(z0=dr_in.z1 + 3.02 * M, L=2e-3, name="laser", w=0.05, h=0.05)
    return Beamline([fourK, fortyK, bb, rc_in, rc_out, spa_in, spa_out, dr_in, laser])


def standard_ics(n, seed, sigma_perp=39.5):
    """CeNTREX-shaped initial conditions from a seeded NumPy generator (host side, tests only)."""
    rng = np.random.default_rng(seed)
    ic = np.empty((6, n))
    th, rr = rng.uniform(0, 2 * np.pi, n), np.sqrt(rng.uniform(0, 1, n)) * 0.01
    ic[0], ic[1], ic[2] = rr * np.cos(th), rr * np.sin(th), 0.25 * 0.0254
    ic[3], ic[4], ic[5] = rng.normal(0, sigma_perp, n), rng.normal(0, sigma_perp, n), rng.normal(184, 16, n)
    return ic
