"""The walk kernel's FP32 fate filter on the device (csrc/cmt_device.cuh: filter_fate, draw_f32).

The filter decides a fate in single precision only when the molecule misses every edge by more than
a rigorous error bound; everything else takes the binary64 path of the reference.  So results must be
IDENTICAL with the filter on (default, whenever no final rows are requested) and off
(cmt_debug_flags bit 1): fates, Counter, saved indices and the work counters -- on the golden
fixtures, against the oracle, on molecules aimed at the edges, on hostile inputs, for both sources,
at sizes up to 2e7 molecules."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle
from tests import filter_model as fm
from tests.beamlines import apertures_beamline, lens_beamline, lens_table, spa_beamline, standard_ics


@pytest.fixture(scope="module")
def torch_cuda(cuda_lib):
    import torch

    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch


def propagator(cuda_lib, elements, flags=0, math="exact"):
    """A Propagator whose beamline handle was created under the given debug flags."""
    from trajectories import _engine as eng

    flat = eng.flatten(elements)
    old = cuda_lib.cmt_debug_flags(flags)
    try:
        prop = eng.Propagator(flat, 0, math=math)
        prop.dev = eng.DeviceBeamline(flat, 0, math)     # bypass the handle cache: flags are read at creation
    finally:
        cuda_lib.cmt_debug_flags(old)
    prop.reset()
    return prop


def fates_only(torch, prop, ic, **kw):
    prop.reset()
    res = prop.propagate_ic(torch.from_numpy(np.ascontiguousarray(ic)).cuda(), want_fate=True, want_final=False, **kw)
    torch.cuda.synchronize()
    return res.fate.cpu().numpy(), res.counters.cpu().numpy().copy(), res.work.cpu().numpy().copy(), res


def check_vs_oracle(torch, cuda_lib, beamline, ic, min_filtered=None):
    want = oracle.propagate(beamline.elements, ic)
    fate, counters, work, _ = fates_only(torch, propagator(cuda_lib, beamline.elements), ic)
    np.testing.assert_array_equal(fate, want["fate"])
    np.testing.assert_array_equal(counters, want["counters"])
    np.testing.assert_array_equal(work[:3], want["work"])            # rows, RK steps, out-of-range evaluations
    if min_filtered is not None:
        assert work[5] >= min_filtered * ic.shape[1], "the filter did not run"
    return work


@pytest.mark.parametrize("name", ["std_seed0", "std_seed1", "std_seed2", "lens_biased", "edges"])
def test_golden_fixtures_with_the_filter(torch_cuda, cuda_lib, golden_dir, name):
    g = np.load(golden_dir / f"{name}.npz")
    for prefix, bl in (("lens", lens_beamline((g["table_r"], g["table_a"]))), ("ap", apertures_beamline())):
        if f"{prefix}_fate" not in g:
            continue
        fate, counters, work, _ = fates_only(torch_cuda, propagator(cuda_lib, bl.elements), g["ic"])
        np.testing.assert_array_equal(fate, g[f"{prefix}_fate"])
        np.testing.assert_array_equal(counters, np.bincount(g[f"{prefix}_fate"], minlength=len(counters)))
        assert work[0] + work[1] == (g[f"{prefix}_n_rows"] - 1).sum()
        if name.startswith("std"):
            assert work[5] > 0.99 * g["ic"].shape[1]


def test_golden_spa_with_the_filter(torch_cuda, cuda_lib, golden_dir):
    g = np.load(golden_dir / "spa.npz")
    fate, counters, work, _ = fates_only(torch_cuda, propagator(cuda_lib, spa_beamline().elements), g["ic"])
    np.testing.assert_array_equal(fate, g["spa_fate"])
    assert work[0] == (g["spa_n_rows"] - 1).sum()


@pytest.mark.parametrize("n,seed,sigma", [(300000, 31, 39.5), (60000, 32, 4.0), (257, 33, 4.0), (1, 34, 4.0), (33, 35, 1.0), (64, 36, 39.5)])
def test_oracle_all_beamlines(torch_cuda, cuda_lib, n, seed, sigma):
    ic = standard_ics(n, seed, sigma)
    check_vs_oracle(torch_cuda, cuda_lib, lens_beamline(lens_table()), ic, 0.9 if sigma > 30 and n > 1000 else None)
    check_vs_oracle(torch_cuda, cuda_lib, apertures_beamline(), ic, 0.99 if n > 1000 else None)
    check_vs_oracle(torch_cuda, cuda_lib, spa_beamline(), ic, 0.99 if n > 1000 else None)


@pytest.mark.parametrize("scale", [1e-4, 1e-6, 1e-7, 1e-9])
def test_molecules_aimed_at_the_edges(torch_cuda, cuda_lib, scale):
    """Every plane of every beamline approached to within `scale` of its size: the filter must hand
    the close calls to binary64 and never decide one wrongly."""
    rng = np.random.default_rng(int(-np.log10(scale)))
    for bl in (lens_beamline(lens_table()), apertures_beamline(), spa_beamline()):
        flat = oracle.flatten(bl.elements)
        ic = fm.aimed_ics(flat, 300000, rng, scale, standard_ics(300000, 17))
        check_vs_oracle(torch_cuda, cuda_lib, bl, ic)


def test_hostile_inputs(torch_cuda, cuda_lib):
    base = standard_ics(64, 9, 3.0)
    ic = np.repeat(base, 12, axis=1)
    k = np.arange(ic.shape[1]) % 12
    ic[5, k == 1] = 0.0
    ic[5, k == 2] *= -1
    ic[0, k == 3] = np.nan
    ic[4, k == 4] = np.inf
    ic[5, k == 5] = 1e-300
    ic[5, k == 6] = 1e300
    ic[0, k == 7] = 1e200
    ic[2, k == 8] = lens_beamline(lens_table()).elements[0].z0
    ic[3, k == 9] = -0.0
    ic[1, k == 10] = 5e-324
    ic[5, k == 11] = np.nan
    for bl in (lens_beamline(lens_table()), apertures_beamline(), spa_beamline()):
        check_vs_oracle(torch_cuda, cuda_lib, bl, ic)
    # single-precision overflow / underflow territory
    ic = np.repeat(base, 8, axis=1)
    k = np.arange(ic.shape[1]) % 8
    ic[0, k == 1] = 1e-42
    ic[3, k == 2] = 1e25
    ic[5, k == 3] = 1e-20
    ic[5, k == 4] = 1e25
    ic[2, k == 5] = -1e20
    ic[4, k == 6] = -1e38
    ic[1, k == 7] = 3e38
    for bl in (lens_beamline(lens_table()), apertures_beamline()):
        check_vs_oracle(torch_cuda, cuda_lib, bl, ic)


@pytest.mark.parametrize("n", [1, 2, 63, 127, 128, 129, 100001])
def test_pair_mode_edges_and_unaligned_inputs(torch_cuda, cuda_lib, n):
    """Two molecules per thread: odd counts (the last thread holds one molecule), counts around a 128-molecule
    tile, and initial conditions that cannot be read as aligned 16-byte pairs (odd leading dimension, a view that
    starts at an odd column), which take the one-molecule form.  Fate by fate against the oracle."""
    torch = torch_cuda
    bl = lens_beamline(lens_table())
    ic = standard_ics(n + 3, 50 + n % 7, 12.0)
    want = oracle.propagate(bl.elements, ic)
    prop = propagator(cuda_lib, bl.elements)
    whole = torch.from_numpy(ic).cuda()                       # leading dimension n + 3
    for lo, hi in ((0, n), (1, n + 1), (2, n + 2), (0, n + 3)):
        prop.reset()
        res = prop.propagate_ic(whole[:, lo:hi], want_fate=True)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(res.fate.cpu().numpy(), want["fate"][lo:hi])
        np.testing.assert_array_equal(res.counters.cpu().numpy(), np.bincount(want["fate"][lo:hi], minlength=len(want["counters"])))
    even = torch.from_numpy(np.ascontiguousarray(ic[:, : n + (n % 2)])).cuda()     # aligned, even leading dimension
    prop.reset()
    res = prop.propagate_ic(even[:, :n], want_fate=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(res.fate.cpu().numpy(), want["fate"][:n])


def test_unusual_geometry(torch_cuda, cuda_lib):
    """Closed aperture (d = 0: thresholds the filter cannot use), offset rectangles, a field plate first,
    elements after a lens, more planes than the filter table holds."""
    from trajectories.beamline import Beamline
    from trajectories.beamline_elements.apertures import CircularAperture, FieldPlates, RectangularAperture
    from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens, make_interpolator

    table = lens_table()
    ic = standard_ics(100000, 41, 6.0)

    def lens(name, z0):
        return ElectrostaticLens(name=name, z0=z0, L=0.1, a_interp=make_interpolator(*table))

    cases = [
        [CircularAperture(z0=0.1, L=0.01, d=0.0, name="closed")],
        [],
        [FieldPlates(z0=0.05, L=0.3, w=0.01, name="plates"), RectangularAperture(z0=0.5, L=0.01, w=0.02, h=0.01, x0=0.003, y0=-0.002, name="rect"),
         CircularAperture(z0=0.7, L=0.01, d=0.05, name="c")],
        [RectangularAperture(z0=0.05, L=0.01, w=0.012, h=0.014, name="rect"), lens("l", 0.2), CircularAperture(z0=0.5, L=0.01, d=0.02, name="after")],
        [CircularAperture(z0=0.02 + 0.01 * i, L=0.004, d=0.03 + 0.002 * i, name=f"a{i}") for i in range(20)],
    ]
    for elements in cases:
        check_vs_oracle(torch_cuda, cuda_lib, Beamline(elements), ic)


def test_saved_indices_with_the_filter(torch_cuda, cuda_lib):
    bl = lens_beamline(lens_table())
    ic = standard_ics(200000, 21, 20.0)
    want = oracle.propagate(bl.elements, ic)
    names = want["fate_names"]
    mask = (1 << names.index("40K shield")) | (1 << names.index("Detected")) | (1 << names.index("Lens entrance"))
    _, _, work, res = fates_only(torch_cuda, propagator(cuda_lib, bl.elements), ic, first_index=5_000_000_000, save_mask=mask)
    expect = np.nonzero(np.isin(want["fate"], [names.index(k) for k in ("40K shield", "Detected", "Lens entrance")]))[0]
    np.testing.assert_array_equal(res.saved_index.cpu().numpy(), expect + 5_000_000_000)
    assert work[5] > 0.9 * ic.shape[1]


@pytest.mark.parametrize("math", ["exact", "contracted"])
def test_filter_on_equals_filter_off_at_scale(torch_cuda, cuda_lib, math):
    """2e7 molecules per beamline and source: identical fates and counters with and without the filter."""
    from trajectories import _engine as eng
    from trajectories.distributions import (CeNTREXPositionDistribution, CeNTREXVelocityDistribution,
                                            GaussianPositionDistribution)

    torch = torch_cuda
    n = 20_000_000 if math == "exact" else 5_000_000
    cases = [(lens_beamline(lens_table()), CeNTREXPositionDistribution(), 0.99),
             (apertures_beamline(), CeNTREXPositionDistribution(), 0.999),
             (spa_beamline(), GaussianPositionDistribution(), 0.999)]
    for bl, xdist, min_filtered in cases:
        src = eng.make_source(CeNTREXVelocityDistribution(), xdist)
        on, off = propagator(cuda_lib, bl.elements, 0, math), propagator(cuda_lib, bl.elements, 2, math)
        slow = propagator(cuda_lib, bl.elements, 4, math)      # per-molecule tolerances only
        single = propagator(cuda_lib, bl.elements, 8, math)    # constant thresholds, one molecule per thread (no pairs)
        # Philox source
        a = on.propagate_philox(src, 99, 1 << 35, n, want_fate=True)
        fa, ca, wa = a.fate.clone(), a.counters.clone(), a.work.clone()
        b = off.propagate_philox(src, 99, 1 << 35, n, want_fate=True)
        c = slow.propagate_philox(src, 99, 1 << 35, n, want_fate=True)
        d = single.propagate_philox(src, 99, 1 << 35, n, want_fate=True)
        torch.cuda.synchronize()
        assert torch.equal(fa, d.fate) and torch.equal(ca, d.counters) and torch.equal(wa[:6], d.work[:6])
        assert torch.equal(fa, b.fate) and torch.equal(ca, b.counters)
        assert torch.equal(fa, c.fate) and torch.equal(ca, c.counters)
        assert torch.equal(wa[:5], b.work[:5]) and torch.equal(wa[:5], c.work[:5])
        assert int(wa[5]) >= (min_filtered - 0.01) * n and int(b.work[5]) == 0
        assert int(c.work[5]) >= min_filtered * n and int(c.work[5]) >= int(wa[5])
        # replayed initial conditions
        ic = on.draw(src, 99, 1 << 35, n)
        on.reset(), off.reset(), slow.reset(), single.reset()
        a = on.propagate_ic(ic, want_fate=True)
        b = off.propagate_ic(ic, want_fate=True)
        c = slow.propagate_ic(ic, want_fate=True)
        d = single.propagate_ic(ic, want_fate=True)
        torch.cuda.synchronize()
        assert torch.equal(a.fate, d.fate) and torch.equal(a.counters, d.counters) and torch.equal(a.work[:6], d.work[:6])
        assert torch.equal(a.fate, b.fate) and torch.equal(a.counters, b.counters)
        assert torch.equal(a.fate, c.fate) and torch.equal(a.counters, c.counters)
        assert torch.equal(a.fate, fa)                     # and the same as the Philox run
        assert torch.equal(a.work[:5], b.work[:5]) and torch.equal(a.work[:5], c.work[:5])
        assert int(a.work[5]) >= (min_filtered - 0.005) * n


@pytest.mark.parametrize("flags", [0, 4, 8])
def test_both_filter_forms_on_aimed_and_hostile_inputs(torch_cuda, cuda_lib, flags):
    """The constant-threshold form two molecules per thread (default) and one per thread (flag 8), and the
    per-molecule-tolerance form (flag 4), separately."""
    rng = np.random.default_rng(77 + flags)
    base = standard_ics(64, 9, 3.0)
    hostile = np.repeat(base, 10, axis=1)
    k = np.arange(hostile.shape[1]) % 10
    hostile[5, k == 1] = 0.0
    hostile[5, k == 2] *= -1
    hostile[0, k == 3] = np.nan
    hostile[4, k == 4] = np.inf
    hostile[5, k == 5] = 1e-300
    hostile[0, k == 7] = 1e200
    hostile[2, k == 8] = -3.0
    hostile[5, k == 9] = 20.0
    for bl in (lens_beamline(lens_table()), apertures_beamline(), spa_beamline()):
        flat = oracle.flatten(bl.elements)
        for ic in (fm.aimed_ics(flat, 200000, rng, 1e-6, standard_ics(200000, 19)),
                   fm.aimed_ics(flat, 200000, rng, 1e-8, standard_ics(200000, 20)), hostile):
            want = oracle.propagate(bl.elements, ic)
            fate, counters, work, _ = fates_only(torch_cuda, propagator(cuda_lib, bl.elements, flags), ic)
            np.testing.assert_array_equal(fate, want["fate"])
            np.testing.assert_array_equal(work[:3], want["work"])


def test_single_precision_source_against_the_model(torch_cuda, cuda_lib):
    """The filter's view of the Philox source: fates of a Philox run equal the oracle's on the oracle's
    own binary64 samples (which the CPU model bounds the single-precision transforms against)."""
    from trajectories import _engine as eng
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution

    bl = lens_beamline(lens_table())
    vdist, xdist = CeNTREXVelocityDistribution(), CeNTREXPositionDistribution()
    n, seed, first = 1_000_000, 0xDEADBEEFCAFE, (1 << 33) + 5
    want = oracle.run(bl.elements, oracle.make_source(vdist, xdist), seed=seed, first=first, n=n)
    prop = propagator(cuda_lib, bl.elements)
    res = prop.propagate_philox(eng.make_source(vdist, xdist), seed, first, n)
    got = res.counters.cpu().numpy()
    assert got.sum() == n
    assert np.abs(got - want["counters"]).sum() <= 2      # CUDA libm vs glibc transforms on the binary64 path
    assert int(res.work[5]) > 0.99 * n
