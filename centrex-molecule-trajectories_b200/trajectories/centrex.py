"""The CeNTREX beamlines of the reference's example scripts as library functions.

lens_beamline      examples/lens_simulation_beamline.py:21-72   (BASELINE.json configs[1], [2], [4])
apertures_beamline the same without the lens                     (configs[0])
spa_beamline       examples/SPA/SPA_distributions.py:21-84       (configs[3])
"""
from __future__ import annotations

from . import _tlf
from .beamline import Beamline
from .beamline_elements.apertures import CircularAperture, FieldPlates, RectangularAperture
from .beamline_elements.electrostatic_lens import ElectrostaticLens, make_interpolator

M_PER_IN = 0.0254


def lens_table(J=2, mJ=0, V=27.6e3, d=1.75 * 0.0254, mass=(204.38 + 19.00) * 1.67e-27):
    """(r, a_r) table for one state / voltage from the built-in Stark model."""
    return _tlf.lens_acceleration_table(d, V, mass, J, mJ)


def _front():
    m = M_PER_IN
    fourK = CircularAperture(z0=1.7 * m, L=0.25 * m, d=1 * m, name="4K shield")
    fortyK = CircularAperture(z0=fourK.z1 + 1.25 * m, L=0.25 * m, d=1 * m, name="40K shield")
    bb = CircularAperture(z0=fortyK.z1 + 2.5 * m, L=0.75 * m, d=4 * m, name="BB exit")
    return fourK, fortyK, bb


def lens_beamline(table=None, V=27.6e3, state=None):
    m = M_PER_IN
    fourK, fortyK, bb = _front()
    lens = ElectrostaticLens(z0=bb.z1 + 33 * m, L=0.6, name="ES lens", V=V)
    if state is not None:
        lens.state = state
    if table is not None:
        lens.a_interp = make_interpolator(*table)
    fp = FieldPlates(z0=2.43, L=3.0, w=0.02, name="Field plates")
    dr = RectangularAperture(z0=fp.z1 + 39.9 * m, L=0.25 * m, name="DR aperture", w=0.018, h=0.03)
    return Beamline([fourK, fortyK, bb, lens, fp, dr])


def apertures_beamline():
    m = M_PER_IN
    fourK, fortyK, bb = _front()
    fp = FieldPlates(z0=2.43, L=3.0, w=0.02, name="Field plates")
    dr = RectangularAperture(z0=fp.z1 + 39.9 * m, L=0.25 * m, name="DR aperture", w=0.018, h=0.03)
    return Beamline([fourK, fortyK, bb, fp, dr])


def spa_beamline():
    m = M_PER_IN
    fourK, fortyK, bb = _front()
    rc_in = CircularAperture(z0=17.36 * m, L=0.125 * m, d=8e-3, name="RC entrance")
    rc_out = CircularAperture(z0=(17.36 + 9) * m, L=0.125 * m, d=8e-3, name="RC exit")
    spa_in = CircularAperture(z0=bb.z1 + 20.5 * m, L=0.375 * m, d=1.75 * m, name="SPA entrance")
    spa_out = CircularAperture(z0=spa_in.z1 + 9.625 * m, L=0.375 * m, d=1.75 * m, name="SPA exit")
    dr_in = CircularAperture(z0=(35.37 + 11) * m, L=0.125 * m, d=150e-3, name="DR entrance")
    laser = RectangularAperture(z0=dr_in.z1 + 3.02 * m, L=2e-3, name="laser", w=0.05, h=0.05)
    return Beamline([fourK, fortyK, bb, rc_in, rc_out, spa_in, spa_out, dr_in, laser])
