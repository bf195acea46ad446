"""Stark shift of a TlF X-state level versus electric field (host side).

Mirror of the reference's `stark_potential(state, Ezs)` (stark_potential.py:9-67):
Ezs in V/cm, result in joule.  The reference needs the external, unpinned
`centrex_TlF` package; when it is importable the full Hamiltonian is used
exactly as there (minus the blocking plt.show() of lines 63-65), otherwise the
build's own X-state Hamiltonian (`_tlf_full.py`: rotation, Stark, hyperfine and
Zeeman terms in the same 196-state uncoupled basis, the state followed by the
same overlap re-ordering) supplies the curve.  Parity at this boundary is
unpinned either way (nothing in the reference pins centrex_TlF's numbers), and
the table the curve feeds (`ElectrostaticLens.a_interp`) can always be injected
by the caller.  `MODEL = "rigid"` selects the bare rigid rotor of `_tlf.py`
(no hyperfine structure; within 1e-5 of the full curve at lens fields).
"""
from __future__ import annotations

import numpy as np

from . import _tlf, _tlf_full

_CURVES: dict = {}
MODEL = "full"      # "full": _tlf_full (the reference's physics); "rigid": _tlf.rigid_rotor_stark_joule

try:  # pragma: no cover - not installed in the build image
    from centrex_TlF.hamiltonian import generate_uncoupled_hamiltonian_X  # type: ignore  # noqa: F401
    from centrex_TlF.states import State, UncoupledBasisState  # type: ignore

    HAVE_CENTREX_TLF = True      # the real package (the state-only shim has no Hamiltonian)
except Exception:  # ImportError or a broken install
    from ._states import State, UncoupledBasisState

    HAVE_CENTREX_TLF = False

__all__ = ["stark_potential", "reorder_evecs", "State", "UncoupledBasisState"]


def default_lens_state():
    """|J=2, mJ=0> with the nuclear spins of electrostatic_lens.py:33-43."""
    return 1 * UncoupledBasisState(J=2, mJ=0, I1=1 / 2, m1=1 / 2, I2=1 / 2, m2=-1 / 2, Omega=0, P=+1,
                                   electronic_state="X")


def state_spin_projections(state):
    """(m1, m2) of the state's largest component; the nuclear spins of electrostatic_lens.py:33-43 for a bare
    (J, mJ) pair."""
    if isinstance(state, (tuple, list)):
        return (float(state[2]), float(state[3])) if len(state) >= 4 else (0.5, -0.5)
    c = state.find_largest_component()
    return float(c.m1), float(c.m2)


def state_quantum_numbers(state):
    """(J, mJ) of the state's largest component (electrostatic_lens.py:176-177)."""
    if isinstance(state, (tuple, list)) and len(state) in (2, 4):
        return int(state[0]), int(state[1])
    c = state.find_largest_component()
    return int(c.J), int(c.mJ)


def reorder_evecs(V_in, E_in, V_ref):
    """Reorder eigenpairs to follow the reference set by maximum overlap (stark_potential.py:78-88)."""
    overlap = np.abs(np.conj(V_in.T) @ V_ref)
    order = np.argsort(np.argmax(overlap, axis=1))
    return E_in[order], V_in[:, order]


def _stark_centrex_tlf(state, Ezs):  # pragma: no cover - needs the external package
    from scipy.constants import hbar
    from centrex_TlF.hamiltonian import (generate_uncoupled_hamiltonian_X,
                                         generate_uncoupled_hamiltonian_X_function)
    from centrex_TlF.states import generate_uncoupled_states_ground
    from centrex_TlF.states.utils import find_closest_vector_idx

    QN = generate_uncoupled_states_ground([0, 1, 2, 3, 4, 5, 6])
    H = generate_uncoupled_hamiltonian_X_function(generate_uncoupled_hamiltonian_X(QN))
    B = np.array((0, 0, 0.0001))
    _, V_ref = np.linalg.eigh(H(np.array((0, 0, 100)), B))
    V_ref_0 = V_ref
    energies = np.zeros((len(Ezs), V_ref.shape[0]))
    for i, Ez in enumerate(Ezs):
        D, V = np.linalg.eigh(H(np.array((0, 0, Ez)), B))
        D, V = reorder_evecs(V, D, V_ref)
        energies[i] = D
        V_ref = V
    idx = find_closest_vector_idx(state.state_vector(QN), V_ref_0)
    return energies[:, idx] * hbar


def stark_potential(state, Ezs):
    """Stark potential (J) of `state` at field magnitudes `Ezs` (V/cm)."""
    Ezs = np.asarray(Ezs, dtype=np.float64)
    if HAVE_CENTREX_TLF:  # pragma: no cover
        return _stark_centrex_tlf(state, Ezs)
    J, mJ = state_quantum_numbers(state)
    if MODEL == "rigid":
        return _tlf.rigid_rotor_stark_joule(J, mJ, Ezs)
    m1, m2 = state_spin_projections(state)
    # a sweep or a loop of runs asks for the same curve again and again (25 ms each: 222 small eigh calls)
    key = (J, mJ, m1, m2, Ezs.shape, Ezs.tobytes())
    hit = _CURVES.get(key)
    if hit is None:
        if len(_CURVES) >= 256:
            _CURVES.clear()
        hit = _CURVES[key] = _tlf_full.stark_joule(J, mJ, m1, m2, Ezs)
    return hit.copy()
