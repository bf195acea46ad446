"""Built-in Stark-shift model for TlF X-state rotational levels (host side).

The reference obtains the Stark curve from the external, unpinned `centrex_TlF`
package (stark_potential.py:9-67): it diagonalises the full X-state Hamiltonian
(rotation + Stark + hyperfine + Zeeman) for J = 0..6 in the uncoupled basis and
follows one adiabatic state.  That package is not available to this build and
no reference test pins its output, so PARITY IS UNPINNED at this boundary.

This module supplies the physics that dominates the curve: a rigid rotor in a
static field, H = B J(J+1) - mu E cos(theta), which is block diagonal in mJ.
For |J, mJ> the block is tridiagonal in J' = |mJ|..J_max with
<J,m|cos(theta)|J+1,m> = sqrt(((J+1)^2 - m^2) / ((2J+1)(2J+3))).
The hyperfine and Zeeman terms the reference includes are ~1e-5 of the Stark
shift at lens fields.  J_max = 6 mirrors the reference's basis truncation
(stark_potential.py:24).  When the real `centrex_TlF` is importable,
`trajectories.stark_potential.stark_potential` uses it instead.

The module is deliberately import-free apart from numpy so that test tooling
can load it by file path.
"""
from __future__ import annotations

import numpy as np

# TlF X(v=0) constants (values as used by the CeNTREX collaboration's code).
B_ROT_HZ = 6689920000.0  # rotational constant, Hz
# 4.2282 D in Hz/(V/cm): D[e a0] * a0[cm] / h[eV s]
D_TLF_HZ_PER_V_CM = 4.2282 * 0.393430307 * 5.291772e-9 / 4.135667e-15
H_PLANCK = 6.62607015e-34  # J s
J_MAX = 6


def rigid_rotor_energies_hz(J: int, mJ: int, Ez_V_per_cm, j_max: int = J_MAX) -> np.ndarray:
    """Energy (Hz) of the adiabatic |J, mJ> rigid-rotor state at each field value."""
    m = abs(int(mJ))
    J = int(J)
    if J < m or J > j_max:
        raise ValueError(f"need |mJ| <= J <= {j_max}, got J={J}, mJ={mJ}")
    Js = np.arange(m, j_max + 1, dtype=np.float64)
    diag = B_ROT_HZ * Js * (Js + 1.0)
    jj = Js[:-1]
    off = np.sqrt(((jj + 1.0) ** 2 - m * m) / ((2.0 * jj + 1.0) * (2.0 * jj + 3.0)))
    Ez = np.atleast_1d(np.asarray(Ez_V_per_cm, dtype=np.float64))
    k = J - m  # no crossings inside one mJ block: the k-th eigenvalue stays the k-th
    H0 = np.diag(diag)
    C = np.diag(off, 1) + np.diag(off, -1)
    # one stacked call: LAPACK diagonalises the matrices one after the other, as a Python loop over them would
    H = H0[None, :, :] - (D_TLF_HZ_PER_V_CM * Ez.reshape(-1))[:, None, None] * C[None, :, :]
    return np.linalg.eigvalsh(H)[:, k].reshape(Ez.shape)


def rigid_rotor_stark_joule(J: int, mJ: int, Ez_V_per_cm) -> np.ndarray:
    """Stark potential in joule, the unit stark_potential.py:57-61 returns."""
    return rigid_rotor_energies_hz(J, mJ, Ez_V_per_cm) * H_PLANCK


def lens_acceleration_table(d: float, V: float, mass: float, J: int, mJ: int, stark=None):
    """(r_values, a_values) exactly as electrostatic_lens.py:194-206 builds them.

    Quirks kept on purpose: the grid has int(round(d/2/1e-4)) points spanning
    [0, 1.01 d/2] (true spacing != dr) while np.gradient is given the nominal
    dr = 1e-4; the end points use one-sided differences so a_r(0) != 0.
    `stark(Ez_V_per_cm) -> joule` defaults to the rigid-rotor model above.
    """
    dr = 1e-4
    r_values = np.linspace(0, d / 2 * 1.01, int(np.round(d / 2 / dr)))
    E_values = 2 * V / ((d / 2) ** 2) * r_values  # V/m
    if stark is None:
        V_stark = rigid_rotor_stark_joule(J, mJ, E_values / 100)
    else:
        V_stark = np.asarray(stark(E_values / 100), dtype=np.float64)
    a_values = -np.gradient(V_stark, dr) / mass
    return r_values, a_values
