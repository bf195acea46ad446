"""The error model of the walk kernel's FP32 fate filter, checked on the CPU: a NumPy float32
restatement of the device code (tests/filter_model.py) must never decide a fate the oracle
disagrees with -- on CeNTREX-shaped inputs, on molecules aimed at the aperture edges to within
1e-8 of their size, and with the device source's single-precision transforms perturbed by the
documented worst-case error of the fast intrinsics.  The GPU tests (tests/test_gpu_filter.py)
check the device code itself the same way."""
import numpy as np
import pytest

from oracle import oracle
from tests import filter_model as fm
from tests.beamlines import apertures_beamline, lens_beamline, lens_table, spa_beamline, standard_ics


def beamlines():
    return {"lens": lens_beamline(lens_table()), "apertures": apertures_beamline(), "spa": spa_beamline()}


def judge(flat, ic, q=None):
    want = oracle.propagate(flat, ic)
    planes, covers_all = fm.filter_planes(flat)
    fate, rows = fm.filter_fate(planes, covers_all, flat.fate_detected, fm.filter_input(ic) if q is None else q)
    decided = fate >= 0
    assert not (decided & (fate != want["fate"])).any()
    assert not (decided & (rows != want["n_rows"] - 1)).any()
    return decided.mean(), want


@pytest.mark.parametrize("name", ["lens", "apertures", "spa"])
def test_standard_inputs(name):
    flat = oracle.flatten(beamlines()[name].elements)
    frac, want = judge(flat, standard_ics(400000, 3))
    # everything but the lens survivors (0.54 %) is decided
    assert frac > (0.993 if name == "lens" else 0.9999)


@pytest.mark.parametrize("name", ["lens", "apertures", "spa"])
@pytest.mark.parametrize("scale", [1e-4, 1e-6, 1e-8])
def test_molecules_aimed_at_the_edges(name, scale):
    flat = oracle.flatten(beamlines()[name].elements)
    rng = np.random.default_rng(int(-np.log10(scale)))
    ic = fm.aimed_ics(flat, 150000, rng, scale, standard_ics(150000, 17))
    judge(flat, ic)


def test_golden_edge_fixture(golden_dir):
    g = np.load(golden_dir / "edges.npz")
    for bl in (lens_beamline((g["table_r"], g["table_a"])), apertures_beamline()):
        judge(oracle.flatten(bl.elements), g["ic"])


def test_hostile_inputs_are_left_to_binary64():
    flat = oracle.flatten(beamlines()["lens"].elements)
    ic = np.repeat(standard_ics(64, 9, 3.0), 8, axis=1)
    k = np.arange(ic.shape[1]) % 8
    ic[5, k == 1] = 0.0
    ic[5, k == 2] *= -1
    ic[0, k == 3] = np.nan
    ic[4, k == 4] = np.inf
    ic[5, k == 5] = 1e-300
    ic[5, k == 6] = 1e300
    ic[0, k == 7] = 1e200
    judge(flat, ic)


def test_unusable_thresholds_switch_the_filter_off():
    from trajectories.beamline import Beamline
    from trajectories.beamline_elements import CircularAperture

    flat = oracle.flatten(Beamline([CircularAperture(z0=0.1, L=0.01, d=0.0, name="closed")]).elements)
    assert fm.filter_planes(flat) == ([], False)


@pytest.mark.parametrize("pos", ["disc", "gauss"])
def test_single_precision_source_stays_within_its_bounds(pos):
    from trajectories.distributions import (CeNTREXPositionDistribution, CeNTREXVelocityDistribution,
                                            GaussianPositionDistribution)

    xdist = CeNTREXPositionDistribution() if pos == "disc" else GaussianPositionDistribution()
    src = oracle.make_source(CeNTREXVelocityDistribution(), xdist)
    rng = np.random.default_rng(1)
    n, seed, first = 400000, 0xDEADBEEFCAFE, (1 << 33) + 5
    ic = oracle.draw(src, seed, first, n)
    for noise in (None, rng):
        q = fm.draw_f32(src, seed, first, n, noise)
        for k, i, e in (("x0", 0, "ex0"), ("y0", 1, "ey0"), ("vx", 3, "evx"), ("vy", 4, "evy"), ("vz", 5, "evz")):
            ok = np.isfinite(q[k])
            err = np.abs(q[k][ok].astype(np.float64) - ic[i][ok])
            assert (err <= 0.5 * q[e][ok]).all()          # the stated bounds carry a factor two to spare
    # and the whole filter on those samples
    for name, bl in beamlines().items():
        if (name == "spa") != (pos == "gauss"):
            continue
        flat = oracle.flatten(bl.elements)
        frac, _ = judge(flat, ic, fm.draw_f32(src, seed, first, n, rng))
        assert frac > 0.99


# ---- the filter with constant thresholds (quick_fate) ----
def judge_quick(flat, ic, source=None, q=None):
    want = oracle.propagate(flat, ic)
    _, covers_all = fm.filter_planes(flat)
    table = fm.quick_table(flat, source)
    assert table is not None
    fate, rows, guarded = fm.quick_fate(table, covers_all, flat.fate_detected, fm.filter_input(ic) if q is None else q)
    decided = fate >= 0
    assert not (decided & (fate != want["fate"])).any()
    assert not (decided & (rows != want["n_rows"] - 1)).any()
    return decided.mean(), guarded.mean()


@pytest.mark.parametrize("name", ["lens", "apertures", "spa"])
def test_quick_standard_inputs(name):
    flat = oracle.flatten(beamlines()[name].elements)
    frac, guarded = judge_quick(flat, standard_ics(400000, 3))
    assert guarded > 0.999 and frac > (0.993 if name == "lens" else 0.999)


@pytest.mark.parametrize("name", ["lens", "apertures", "spa"])
@pytest.mark.parametrize("scale", [1e-4, 1e-6, 1e-8])
def test_quick_molecules_aimed_at_the_edges(name, scale):
    flat = oracle.flatten(beamlines()[name].elements)
    rng = np.random.default_rng(10 + int(-np.log10(scale)))
    judge_quick(flat, fm.aimed_ics(flat, 150000, rng, scale, standard_ics(150000, 18)))


def test_quick_hostile_and_offset_inputs(golden_dir):
    flat = oracle.flatten(beamlines()["lens"].elements)
    ic = np.repeat(standard_ics(64, 9, 3.0), 10, axis=1)
    k = np.arange(ic.shape[1]) % 10
    ic[5, k == 1] = 0.0
    ic[5, k == 2] *= -1
    ic[0, k == 3] = np.nan
    ic[4, k == 4] = np.inf
    ic[5, k == 5] = 1e-300
    ic[5, k == 6] = 1e300
    ic[0, k == 7] = 1e200
    ic[2, k == 8] = -3.0          # far upstream: outside the z0 guard
    ic[5, k == 9] = 20.0          # slow: gravity matters, outside the velocity guard
    judge_quick(flat, ic)
    g = np.load(golden_dir / "edges.npz")
    for bl in (lens_beamline((g["table_r"], g["table_a"])), apertures_beamline()):
        judge_quick(oracle.flatten(bl.elements), g["ic"])


@pytest.mark.parametrize("pos", ["disc", "gauss"])
def test_quick_with_the_single_precision_source(pos):
    from trajectories.distributions import (CeNTREXPositionDistribution, CeNTREXVelocityDistribution,
                                            GaussianPositionDistribution)

    xdist = CeNTREXPositionDistribution() if pos == "disc" else GaussianPositionDistribution()
    src = oracle.make_source(CeNTREXVelocityDistribution(), xdist)
    rng = np.random.default_rng(2)
    n, seed, first = 400000, 99, 1 << 35
    ic = oracle.draw(src, seed, first, n)
    for name, bl in beamlines().items():
        if (name == "spa") != (pos == "gauss"):
            continue
        flat = oracle.flatten(bl.elements)
        frac, guarded = judge_quick(flat, ic, src, fm.draw_f32(src, seed, first, n, rng))
        assert guarded > 0.99 and frac > 0.985
