"""The build's TlF X-state Hamiltonian (trajectories/_tlf_full.py), the stand-in for the external centrex_TlF package
behind the reference's stark_potential(state, Ezs) (stark_potential.py:9-67).  Parity with centrex_TlF is unpinned
(the package is absent and the reference pins none of its numbers); what is checked here:

* the operator algebra against closed forms (J = 0 and J = 1 hyperfine structure, whose measured intervals are in
  the TlF literature), and the Stark part against the rigid rotor;
* the block-wise, batched state following against a literal transcription of the reference's loop (196 x 196
  `eigh` per field value + `reorder_evecs`);
* how far the lens acceleration table built from it is from the rigid-rotor table."""
import numpy as np
import pytest

from trajectories import _tlf, _tlf_full as F
from trajectories import stark_potential as sp


def test_basis_and_hermiticity():
    QN, H_ff, H_S, H_Z = F.operators()
    assert len(QN) == 196 and QN[0] == (0, 0, -0.5, -0.5) and QN[-1] == (6, 6, 0.5, 0.5)
    for M in (H_ff, H_S, H_Z):
        assert np.array_equal(M, M.T)
    # fields along z conserve mF
    mF = np.array([q[1] + q[2] + q[3] for q in QN])
    H = F.hamiltonian(31000.0, 0.3)
    assert np.all(H[mF[:, None] != mF[None, :]] == 0.0)


def test_field_free_hyperfine_structure():
    c = F.XConstants()
    w = np.linalg.eigvalsh(F.hamiltonian(0.0, 0.0))
    # J = 0: F = 0 and F = 1, split by c4 <I1.I2> = c4 (1/4 - (-3/4))
    j0 = np.sort(w[:4])
    assert np.allclose(j0[:3] - j0[3], c.c4, atol=1e-3) and abs(j0[3] - j0[0] + c.c4) < 1e-3
    # J = 1: the four levels F1 = 1/2 (F = 0, 1) and F1 = 3/2 (F = 1, 2) with their 2F + 1 degeneracies, and the
    # measured intervals of the literature: 22.24 kHz, 175.95 kHz, 14.54 kHz
    j1 = np.sort(w[4:16]) - 2 * c.B_rot
    levels = [j1[0], j1[1:4].mean(), j1[4:7].mean(), j1[7:12].mean()]
    for lo, hi in ((1, 4), (4, 7), (7, 12)):
        assert np.ptp(j1[lo:hi]) < 1e-3
    gaps = np.diff(levels) / 1e3
    assert abs(gaps[0] - 22.24) < 0.05 and abs(gaps[1] - 175.95) < 0.05 and abs(gaps[2] - 14.54) < 0.05
    # centre of gravity of a J manifold is B J (J + 1): the hyperfine terms are traceless
    for J in range(7):
        lo, hi = 4 * J * J, 4 * (J + 1) ** 2
        assert abs(np.sort(w)[lo:hi].mean() - c.B_rot * J * (J + 1)) < 1e-2


def test_spin_rotation_against_the_coupled_formula():
    """Only c1 switched on: E = B J(J+1) + c1/2 [F1(F1+1) - J(J+1) - 3/4], F1 = J +- 1/2, each level 2 (2 F1 + 1) fold."""
    c = F.XConstants(c2=0.0, c3=0.0, c4=0.0)
    w = np.sort(np.linalg.eigvalsh(F.hamiltonian(0.0, 0.0, consts=c)))
    want = []
    for J in range(7):
        for F1 in ([0.5] if J == 0 else [J - 0.5, J + 0.5]):
            e = c.B_rot * J * (J + 1) + c.c1 / 2 * (F1 * (F1 + 1) - J * (J + 1) - 0.75)
            want += [e] * int(2 * (2 * F1 + 1))
    assert np.allclose(w, np.sort(want), rtol=0, atol=1e-3)


def test_stark_part_is_the_rigid_rotor():
    bare = F.XConstants(c1=0, c2=0, c3=0, c4=0, mu_J=0, mu_Tl=0, mu_F=0)
    Ez = np.linspace(0, 70000, 64)
    for J, mJ in ((0, 0), (1, 1), (2, 0), (3, 2), (6, 0)):
        got = F.follow_state(J, mJ, 0.5, -0.5, Ez, consts=bare)
        want = _tlf.rigid_rotor_energies_hz(J, mJ, Ez)
        assert np.max(np.abs(got - want)) < 1e-4 * 1.0 + 1e-14 * np.max(np.abs(want))


def reference_loop(J, mJ, m1, m2, Ezs):
    """stark_potential.py:24-61 transcribed: full matrix, eigh per field value, reorder_evecs, pick by overlap."""
    QN = F.basis()
    B = 1e-4
    _, V_ref = np.linalg.eigh(F.hamiltonian(100.0, B))
    V_ref_0 = V_ref
    energies = np.zeros((len(Ezs), len(QN)))
    for i, Ez in enumerate(Ezs):
        D, V = np.linalg.eigh(F.hamiltonian(Ez, B))
        D, V = sp.reorder_evecs(V, D, V_ref)
        energies[i] = D
        V_ref = V
    vec = np.zeros(len(QN))
    vec[QN.index((J, mJ, m1, m2))] = 1.0
    idx = int(np.argmax(np.abs(V_ref_0.T @ vec)))
    return energies[:, idx]


@pytest.mark.parametrize("state", [(2, 0, 0.5, -0.5), (1, 1, 0.5, 0.5), (3, 1, -0.5, 0.5), (0, 0, 0.5, -0.5)])
def test_blockwise_following_equals_the_reference_loop(state):
    Ezs = np.linspace(0, 62000, 48)          # the lens table's own range: E = 2 V r / (d/2)^2 up to ~ 6e4 V/cm
    got = F.follow_state(*state, Ezs)
    want = reference_loop(*state, Ezs)
    # same adiabatic curve: equal to the accuracy of two different LAPACK problem sizes
    assert np.max(np.abs(got - want)) < 1e-3, np.max(np.abs(got - want))


def test_states_of_one_block_share_its_eigenpairs():
    """A sweep over states asks for the same (mF block, field grid) once per state: the followed energies of the whole
    block are computed once (threads that ask for the same block wait for the first) and every state reads its own
    column -- the same numbers as a computation of its own."""
    from concurrent.futures import ThreadPoolExecutor

    Ez = np.linspace(0.0, 31000.0, 222)
    states = [(0, 0), (1, 0), (2, 0), (3, 0), (1, 1), (2, 1)]
    alone = []
    for J, mJ in states:
        F._FOLLOWED.clear()
        alone.append(F.follow_state(J, mJ, 0.5, -0.5, Ez))
    F._FOLLOWED.clear()
    with ThreadPoolExecutor(max_workers=4) as pool:
        shared = list(pool.map(lambda s: F.follow_state(s[0], s[1], 0.5, -0.5, Ez), states))
    assert len(F._FOLLOWED) == 2                       # mF = 0 and mF = 1, one decomposition each
    for a, b in zip(alone, shared):
        assert np.array_equal(a, b)
    shared[0][:] = 0.0                                 # callers get copies
    assert np.array_equal(F.follow_state(0, 0, 0.5, -0.5, Ez), alone[0])


def test_lens_table_full_vs_rigid():
    """How far the rigid-rotor table (fixtures, bench workload) is from the full-Hamiltonian one the lens builds."""
    mass = (204.38 + 19.00) * 1.67e-27
    d, V = 1.75 * 0.0254, 27.6e3
    r, a_rigid = _tlf.lens_acceleration_table(d, V, mass, 2, 0)
    r2, a_full = _tlf.lens_acceleration_table(d, V, mass, 2, 0, stark=lambda Ez: F.stark_joule(2, 0, 0.5, -0.5, Ez))
    assert np.array_equal(r, r2)
    rel = np.abs(a_full - a_rigid) / np.max(np.abs(a_rigid))
    # largest on the first table points (fields below ~1 kV/cm, where the Stark shift is no larger than the hyperfine
    # structure): 4.8e-4 of the largest acceleration; below 1e-6 over the outer half of the bore
    assert rel.max() < 1e-3 and rel[len(rel) // 2:].max() < 1e-6, (rel.max(), rel[len(rel) // 2:].max())
    # and that is what ElectrostaticLens builds by default
    from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens

    lens = ElectrostaticLens(z0=1.0, L=0.6, name="ES lens")
    x, y = lens.acceleration_table()
    assert sp.MODEL == "full" and np.array_equal(y, a_full) and np.array_equal(x, r)
    assert np.array_equal(sp.stark_potential((2, 0), np.array([0.0, 3e4])), F.stark_joule(2, 0, 0.5, -0.5, np.array([0.0, 3e4])))
