"""TlF X(v=0) Hamiltonian in the uncoupled basis |J, mJ, I1=1/2, m1, I2=1/2, m2>, J = 0..6 (196 states), and the
adiabatic Stark curve of one of its states (host side).

This is what the reference's `stark_potential(state, Ezs)` evaluates through the external `centrex_TlF` package
(stark_potential.py:9-67: `generate_uncoupled_states_ground([0..6])`, `generate_uncoupled_hamiltonian_X`, fields
E = (0, 0, Ez), B = (0, 0, 1e-4 G), eigenvectors re-ordered from one field value to the next by largest overlap,
:78-88).  The package is not in the image and the reference pins none of its numbers, so PARITY IS UNPINNED here;
the model below is the published one:

    H = B J^2                                              rotation
      + c1 I1.J + c2 I2.J + c4 I1.I2                        spin-rotation (Tl, F), scalar spin-spin
      + 5 c3 [3 (I1.J)(I2.J) + 3 (I2.J)(I1.J) - 2 (I1.I2) J^2] / ((2J+3)(2J-1))     tensor spin-spin
      - D E cos(theta)                                     Stark, field along z
      - mu_J (Jz / J) Bz - 2 mu_Tl I1z Bz - 2 mu_F I2z Bz   Zeeman (lifts the +-mF degeneracy, 1e-4 G)

with the constants CeNTREX uses (Wilkening, Ramsey and Larson 1984): B = 6.68992 GHz, c1 = 126.03 kHz,
c2 = 17.89 kHz, c3 = 0.70 kHz, c4 = -13.30 kHz, mu_J = 35 Hz/G, mu_Tl = 1240.5 Hz/G, mu_F = 2003.63 Hz/G,
D = 4.2282 D.  Everything is kept in Hz; the caller converts to joule (the reference: rad/s times hbar).

With fields along z the Hamiltonian conserves mF = mJ + m1 + m2, so it is diagonalised block by block (the largest
block, mF = 0, has 26 states) -- the same eigenvalues and, inside a block, the same eigenvectors as the reference's
196 x 196 `eigh`, without the arbitrary mixing of degenerate vectors of different blocks.  The state is followed
exactly as the reference follows it: reference eigenvectors at 100 V/cm, then for every field value in the order
given `index = argsort(argmax(|V^T V_ref|, axis=1))`, and V_ref becomes the re-ordered set.  The per-field overlaps
are one batched product; only the composition of the permutations runs in a Python loop.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from functools import lru_cache
from typing import Dict, List, Tuple

import numpy as np

H_PLANCK = 6.62607015e-34  # J s
J_MAX = 6


@dataclass(frozen=True)
class XConstants:
    B_rot: float = 6689920000.0
    c1: float = 126030.0
    c2: float = 17890.0
    c3: float = 700.0
    c4: float = -13300.0
    mu_J: float = 35.0
    mu_Tl: float = 1240.5
    mu_F: float = 2003.63
    D_TlF: float = 4.2282 * 0.393430307 * 5.291772e-9 / 4.135667e-15   # Hz / (V/cm)


def basis(j_max: int = J_MAX) -> List[Tuple[int, int, float, float]]:
    """(J, mJ, m1, m2) in the order of centrex_TlF.states.generate_uncoupled_states_ground."""
    out = []
    for J in range(j_max + 1):
        for mJ in range(-J, J + 1):
            for m1 in (-0.5, 0.5):
                for m2 in (-0.5, 0.5):
                    out.append((J, mJ, m1, m2))
    return out


def _ladder(j: float, m: float, up: bool) -> float:
    """<j, m +- 1| J+- |j, m>"""
    mm = m + 1 if up else m - 1
    if abs(mm) > j:
        return 0.0
    return float(np.sqrt(j * (j + 1) - m * mm))


@lru_cache(maxsize=4)
def operators(j_max: int = J_MAX, consts: XConstants = XConstants()):
    """(QN, H_ff, H_stark_z, H_zeeman_z): field-free part in Hz, Stark part in Hz per V/cm, Zeeman part in Hz per G."""
    QN = basis(j_max)
    n = len(QN)
    where = {q: i for i, q in enumerate(QN)}
    Jz, Jp, Jm = (np.zeros((n, n)) for _ in range(3))
    I1z, I1p, I1m = (np.zeros((n, n)) for _ in range(3))
    I2z, I2p, I2m = (np.zeros((n, n)) for _ in range(3))
    J2 = np.zeros((n, n))
    cos = np.zeros((n, n))
    for i, (J, mJ, m1, m2) in enumerate(QN):
        J2[i, i] = J * (J + 1)
        Jz[i, i], I1z[i, i], I2z[i, i] = mJ, m1, m2
        if mJ + 1 <= J:
            Jp[where[(J, mJ + 1, m1, m2)], i] = _ladder(J, mJ, True)
        if mJ - 1 >= -J:
            Jm[where[(J, mJ - 1, m1, m2)], i] = _ladder(J, mJ, False)
        if m1 < 0:
            I1p[where[(J, mJ, 0.5, m2)], i] = 1.0
        else:
            I1m[where[(J, mJ, -0.5, m2)], i] = 1.0
        if m2 < 0:
            I2p[where[(J, mJ, m1, 0.5)], i] = 1.0
        else:
            I2m[where[(J, mJ, m1, -0.5)], i] = 1.0
        if J + 1 <= j_max:
            c = np.sqrt(((J + 1) ** 2 - mJ ** 2) / ((2 * J + 1) * (2 * J + 3)))
            k = where[(J + 1, mJ, m1, m2)]
            cos[k, i] = cos[i, k] = c
    I1J = I1z @ Jz + 0.5 * (I1p @ Jm + I1m @ Jp)
    I2J = I2z @ Jz + 0.5 * (I2p @ Jm + I2m @ Jp)
    I1I2 = I1z @ I2z + 0.5 * (I1p @ I2m + I1m @ I2p)
    Jq = np.array([q[0] for q in QN], dtype=np.float64)
    denom = (2 * Jq + 3) * (2 * Jq - 1)                      # J = 0: -3, and the numerator vanishes there
    tensor = (3 * (I1J @ I2J) + 3 * (I2J @ I1J) - 2 * (I1I2 @ J2)) / denom[None, :]
    H_ff = consts.B_rot * J2 + consts.c1 * I1J + consts.c2 * I2J + consts.c4 * I1I2 + 5 * consts.c3 * tensor
    H_ff = 0.5 * (H_ff + H_ff.T)
    H_S = -consts.D_TlF * cos
    with np.errstate(divide="ignore", invalid="ignore"):
        gJ = np.where(Jq > 0, consts.mu_J / np.where(Jq > 0, Jq, 1.0), 0.0)
    H_Z = -(np.diag(gJ) @ Jz) - 2 * consts.mu_Tl * I1z - 2 * consts.mu_F * I2z
    return QN, H_ff, H_S, H_Z


def hamiltonian(Ez: float, Bz: float = 1e-4, j_max: int = J_MAX, consts: XConstants = XConstants()) -> np.ndarray:
    """The full 4 (j_max + 1)^2-square Hamiltonian in Hz for fields along z (Ez in V/cm, Bz in gauss)."""
    _, H_ff, H_S, H_Z = operators(j_max, consts)
    return H_ff + Ez * H_S + Bz * H_Z


@lru_cache(maxsize=64)
def _block(two_mF: int, j_max: int, consts: XConstants):
    QN, H_ff, H_S, H_Z = operators(j_max, consts)
    idx = np.array([i for i, (J, mJ, m1, m2) in enumerate(QN) if int(round(2 * (mJ + m1 + m2))) == two_mF])
    sub = np.ix_(idx, idx)
    return idx, H_ff[sub], H_S[sub], H_Z[sub]


# The eigenpairs of a block and their re-ordering depend on the block and on the field values only, not on which
# of the block's states is followed: a sweep over states (examples/lens_simulation_different_states.py) asks for the
# same (block, field grid) once per state.  The followed energies of ALL states of a block are therefore computed
# once and kept; threads that ask for the same key wait for the first one instead of repeating its work.
_FOLLOWED: Dict[tuple, tuple] = {}
_FOLLOWED_LOCKS: Dict[tuple, "threading.Lock"] = {}
_FOLLOWED_GUARD = threading.Lock()
_FOLLOWED_MAX = 64


def _followed_block(two_mF: int, flat: np.ndarray, Bz: float, E_ref: float, j_max: int, consts: XConstants):
    """(idx, V_ref0, followed): the block's basis indices, its eigenvectors at E_ref, and followed[i, c] = the energy
    at field i of the adiabatic state that is column c of the reference set (stark_potential.py:27-55)."""
    key = (two_mF, flat.tobytes(), float(Bz), float(E_ref), j_max, consts)
    hit = _FOLLOWED.get(key)
    if hit is not None:
        return hit
    with _FOLLOWED_GUARD:
        lock = _FOLLOWED_LOCKS.setdefault(key, threading.Lock())
    with lock:
        hit = _FOLLOWED.get(key)
        if hit is not None:
            return hit
        idx, H0, HS, HZ = _block(two_mF, j_max, consts)
        base = H0 + Bz * HZ
        _, V_ref0 = np.linalg.eigh(base + E_ref * HS)
        energies, vecs = np.linalg.eigh(base[None, :, :] + flat[:, None, None] * HS[None, :, :])
        # raw overlaps |V_i^T V_{i-1}| (and with the reference set for the first field value), all at once
        prev = np.concatenate([V_ref0[None, :, :], vecs[:-1]], axis=0)
        best = np.argmax(np.abs(np.matmul(vecs.transpose(0, 2, 1), prev)), axis=2)    # [n_E, nb]: column of prev per row
        nb = len(idx)
        followed = np.empty((flat.size, nb))
        inv_perm = np.arange(nb)        # position of raw column c of the previous set inside its re-ordered form
        for i in range(flat.size):
            # reorder_evecs: index = argsort(argmax(overlap with the RE-ORDERED previous set, axis=1))
            index = np.argsort(inv_perm[best[i]], kind="quicksort")
            followed[i] = energies[i, index]
            inv_perm = np.empty(nb, dtype=np.int64)
            inv_perm[index] = np.arange(nb)
        hit = (idx, V_ref0, followed)
        with _FOLLOWED_GUARD:
            if len(_FOLLOWED) >= _FOLLOWED_MAX:
                _FOLLOWED.clear()
                _FOLLOWED_LOCKS.clear()
            _FOLLOWED[key] = hit
        return hit


def follow_state(J: int, mJ: int, m1: float, m2: float, Ez_V_per_cm, Bz: float = 1e-4, E_ref: float = 100.0,
                 j_max: int = J_MAX, consts: XConstants = XConstants()) -> np.ndarray:
    """Energy (Hz) of the adiabatic state that is closest to |J, mJ, m1, m2> at E_ref, at every field of
    `Ez_V_per_cm` in the order given (stark_potential.py:27-61)."""
    J, mJ = int(J), int(mJ)
    if not (0 <= J <= j_max and abs(mJ) <= J and abs(m1) == 0.5 and abs(m2) == 0.5):
        raise ValueError(f"not a state of the basis: J={J}, mJ={mJ}, m1={m1}, m2={m2} (J <= {j_max})")
    QN = operators(j_max, consts)[0]
    Ez = np.asarray(Ez_V_per_cm, dtype=np.float64)
    flat = np.ascontiguousarray(Ez.reshape(-1))
    if flat.size == 0:
        return np.zeros(Ez.shape)
    idx, V_ref0, followed = _followed_block(int(round(2 * (mJ + m1 + m2))), flat, Bz, E_ref, j_max, consts)
    me = int(np.nonzero(idx == QN.index((J, mJ, float(m1), float(m2))))[0][0])
    # find_closest_vector_idx: the reference eigenvector with the largest overlap with the basis state
    col = int(np.argmax(np.abs(V_ref0[me, :])))
    return followed[:, col].reshape(Ez.shape).copy()


def stark_joule(J: int, mJ: int, m1: float, m2: float, Ez_V_per_cm, **kw) -> np.ndarray:
    """The quantity stark_potential.py:57-61 returns: energy of the followed state in joule."""
    return follow_state(J, mJ, m1, m2, Ez_V_per_cm, **kw) * H_PLANCK
