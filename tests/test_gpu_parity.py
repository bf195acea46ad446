"""Parity of the CUDA path (libcmt_b200.so, through its C ABI) with the CPU oracle
and with the golden vectors produced by the unmodified reference.

Tolerances (north_star): fates exact; final state 1e-9 relative on ballistic
segments, 1e-6 through the lens.  What is actually achieved is tighter and is
asserted: the CUDA path rounds every operation like the reference, except that
`dt**2` is an exact square on the GPU while NumPy's scalar power goes through
libm pow() (<= 1 ulp apart in 0.08 % of steps), so final rows agree to ~1e-13.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle
from tests.beamlines import apertures_beamline, lens_beamline, lens_table, spa_beamline, standard_ics

TIGHT = 1e-12


def relerr(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-9))


@pytest.fixture(scope="module")
def torch_cuda(cuda_lib):
    import torch

    assert torch.cuda.is_available(), "GPU tests need a B200"
    assert torch.cuda.get_device_capability(0)[0] == 10
    return torch


def gpu_propagate(torch, beamline, ic, math="exact", **kw):
    from trajectories import _engine as eng

    prop = eng.Propagator(beamline.elements, 0, math=math)
    prop.reset()
    res = prop.propagate_ic(torch.from_numpy(np.ascontiguousarray(ic)).cuda(), want_fate=True, want_final=True, **kw)
    torch.cuda.synchronize()
    return dict(fate=res.fate.cpu().numpy(), fin=res.final.cpu().numpy(), counters=res.counters.cpu().numpy(),
                work=res.work.cpu().numpy(), saved=None if res.saved_index is None else res.saved_index.cpu().numpy(),
                prop=prop)


def check_vs_golden(torch, g, prefix, beamline, tol):
    got = gpu_propagate(torch, beamline, g["ic"])
    np.testing.assert_array_equal(got["fate"], g[f"{prefix}_fate"])          # exact, no edge band needed
    assert relerr(got["fin"], g[f"{prefix}_fin"]) < tol
    np.testing.assert_array_equal(got["counters"], np.bincount(g[f"{prefix}_fate"], minlength=len(got["counters"])))
    assert got["work"][0] + got["work"][1] == (g[f"{prefix}_n_rows"] - 1).sum()   # every row accounted for
    return got


@pytest.mark.parametrize("name", ["std_seed0", "std_seed1", "std_seed2", "lens_biased", "lens_biased_J1m1_20kV", "edges"])
def test_golden_lens_beamline(torch_cuda, golden_dir, name):
    g = np.load(golden_dir / f"{name}.npz")
    got = check_vs_golden(torch_cuda, g, "lens", lens_beamline((g["table_r"], g["table_a"])), TIGHT)
    assert got["work"][2] == 0


@pytest.mark.parametrize("name", ["std_seed0", "std_seed1", "std_seed2", "edges"])
def test_golden_apertures_only(torch_cuda, golden_dir, name):
    g = np.load(golden_dir / f"{name}.npz")
    check_vs_golden(torch_cuda, g, "ap", apertures_beamline(), TIGHT)


def test_golden_spa(torch_cuda, golden_dir):
    g = np.load(golden_dir / "spa.npz")
    check_vs_golden(torch_cuda, g, "spa", spa_beamline(), TIGHT)


def test_golden_trajectory_rows(torch_cuda, golden_dir):
    """Saved trajectories row for row (613 rows for a detected molecule)."""
    from trajectories import _engine as eng

    torch = torch_cuda
    g = np.load(golden_dir / "lens_biased.npz")
    bl = lens_beamline((g["table_r"], g["table_a"]))
    prop = eng.Propagator(bl.elements, 0)
    idx = g["lens_row_idx"]
    ic = torch.from_numpy(np.ascontiguousarray(g["ic"][:, idx])).cuda()
    rows, offs, fate = prop.trajectories(ic)
    n_rows = np.diff(offs)
    off = g["lens_row_off"]
    for k in range(len(idx)):
        want = g["lens_rows"][off[k]:off[k + 1]]
        assert n_rows[k] == want.shape[0] == g["lens_n_rows"][idx[k]]
        assert fate[k] == g["lens_fate"][idx[k]]
        assert relerr(rows[offs[k]:offs[k + 1]], want) < TIGHT
    assert rows.shape == (int(n_rows.sum()), 10)               # compact: no padding rows
    # select path: gather by global index out of the full IC array
    full = torch.from_numpy(np.ascontiguousarray(g["ic"])).cuda()
    sel = torch.from_numpy(idx.astype(np.int64) + 1000).cuda()
    rows2, offs2, fate2 = prop.trajectories(full, select=sel, select_base=1000)
    np.testing.assert_array_equal(offs2, offs)
    np.testing.assert_array_equal(rows2, rows)
    np.testing.assert_array_equal(fate2, fate)
    # batching: a tiny device row budget forces many batches and must give the same rows
    import trajectories._engine as engmod
    old_budget = engmod.ROW_BUDGET_BYTES
    engmod.ROW_BUDGET_BYTES = 2000 * 80
    try:
        rows3, offs3, fate3 = prop.trajectories(full, select=sel, select_base=1000)
    finally:
        engmod.ROW_BUDGET_BYTES = old_budget
    np.testing.assert_array_equal(offs3, offs)
    np.testing.assert_array_equal(rows3, rows)
    # results too large to page-lock come back through the staging buffers into ordinary memory
    old_pinned, old_budget = engmod.PINNED_RESULT_BYTES, engmod.ROW_BUDGET_BYTES
    engmod.PINNED_RESULT_BYTES, engmod.ROW_BUDGET_BYTES = 0, 2000 * 80
    try:
        rows4, offs4, fate4 = prop.trajectories(full, select=sel, select_base=1000)
    finally:
        engmod.PINNED_RESULT_BYTES, engmod.ROW_BUDGET_BYTES = old_pinned, old_budget
    np.testing.assert_array_equal(offs4, offs)
    np.testing.assert_array_equal(rows4, rows)
    # the page-locked block lives as long as any view of it and is recycled afterwards
    view = rows[offs[1]:offs[2]].copy(), rows[offs[1]:offs[2]]
    del rows, rows2, rows3
    for _ in range(3):
        prop.trajectories(full, select=sel, select_base=1000)
    np.testing.assert_array_equal(view[1], view[0])


@pytest.mark.parametrize("n,seed,sigma", [(200000, 11, 39.5), (60000, 12, 4.0), (257, 13, 4.0), (1, 14, 4.0), (33, 15, 1.0)])
def test_oracle_lens_beamline(torch_cuda, n, seed, sigma):
    """Seeded CeNTREX-shaped ICs at sizes the oracle finishes in seconds; ragged sizes included."""
    bl = lens_beamline(lens_table())
    ic = standard_ics(n, seed, sigma)
    want = oracle.propagate(bl.elements, ic)
    got = gpu_propagate(torch_cuda, bl, ic)
    np.testing.assert_array_equal(got["fate"], want["fate"])
    np.testing.assert_array_equal(got["counters"], want["counters"])
    np.testing.assert_array_equal(got["work"][:3], want["work"])
    assert got["work"][4] <= 1e-4 * max(got["work"][1], 1) + 2     # the straight-line RK step almost never falls back
    assert relerr(got["fin"], want["fin"]) < TIGHT
    same = (got["fin"].view(np.int64) == want["fin"].view(np.int64)).mean()
    assert same > 0.995          # almost every value is bit-identical (pow() vs exact square)


def test_oracle_other_states(torch_cuda):
    for (J, mJ, V) in [(0, 0, 30e3), (1, 0, 24e3), (2, 2, 34e3), (3, 1, 20e3)]:
        bl = lens_beamline(lens_table(J=J, mJ=mJ, V=V), V=V)
        ic = standard_ics(20000, 100 + J, 3.0)
        want = oracle.propagate(bl.elements, ic)
        got = gpu_propagate(torch_cuda, bl, ic)
        np.testing.assert_array_equal(got["fate"], want["fate"])
        assert relerr(got["fin"], want["fin"]) < TIGHT


def test_empty_input(torch_cuda):
    bl = lens_beamline(lens_table())
    got = gpu_propagate(torch_cuda, bl, np.empty((6, 0)))
    assert got["counters"].sum() == 0 and got["fate"].shape == (0,)


def test_saved_index(torch_cuda):
    bl = lens_beamline(lens_table())
    ic = standard_ics(50000, 21, 3.0)
    want = oracle.propagate(bl.elements, ic)
    names = want["fate_names"]
    mask = (1 << names.index("Detected")) | (1 << names.index("Inside lens"))
    from trajectories import _engine as eng

    prop = eng.Propagator(bl.elements, 0)
    prop.reset()
    res = prop.propagate_ic(torch_cuda.from_numpy(ic).cuda(), first_index=7_000_000_000, save_mask=mask)
    expect = np.nonzero((want["fate"] == names.index("Detected")) | (want["fate"] == names.index("Inside lens")))[0]
    np.testing.assert_array_equal(res.saved_index.cpu().numpy(), expect + 7_000_000_000)


def test_philox_source(torch_cuda):
    """Device Philox samples equal the oracle's (integer stream bit-exact; the
    Box-Muller/sincos transforms differ by ulps between CUDA and glibc)."""
    from trajectories import _engine as eng
    from trajectories.distributions import (CeNTREXPositionDistribution, CeNTREXVelocityDistribution,
                                            GaussianPositionDistribution)

    bl = lens_beamline(lens_table())
    prop = eng.Propagator(bl.elements, 0)
    for xdist in (CeNTREXPositionDistribution(), GaussianPositionDistribution()):
        vdist = CeNTREXVelocityDistribution()
        src = eng.make_source(vdist, xdist)
        first = (1 << 33) + 5                       # exercises the high counter word
        ic = prop.draw(src, seed=0xDEADBEEFCAFE, first_index=first, n=100000).cpu().numpy()
        want = oracle.draw(oracle.make_source(vdist, xdist), 0xDEADBEEFCAFE, first, 100000)
        scale = np.array([0.01, 0.01, 1, 39.5, 39.5, 184.0])[:, None]
        assert np.max(np.abs(ic - want) / scale) < 1e-13
        idx = torch_cuda.tensor([first + 3, first + 99999, first], dtype=torch_cuda.int64, device="cuda")
        sub = prop.draw(src, seed=0xDEADBEEFCAFE, index=idx).cpu().numpy()
        np.testing.assert_array_equal(sub, ic[:, [3, 99999, 0]])


def test_philox_run_counts(torch_cuda):
    """Whole Philox run vs the oracle's run on the same global indices: counts agree
    except for molecules within rounding of an edge (none expected at this size)."""
    from trajectories import _engine as eng
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution

    bl = lens_beamline(lens_table())
    vdist, xdist = CeNTREXVelocityDistribution(), CeNTREXPositionDistribution()
    n = 400000
    want = oracle.run(bl.elements, oracle.make_source(vdist, xdist), seed=5, first=1000, n=n)
    prop = eng.Propagator(bl.elements, 0)
    prop.reset()
    res = prop.propagate_philox(eng.make_source(vdist, xdist), 5, 1000, n, want_fate=True)
    got = res.counters.cpu().numpy()
    assert got.sum() == n
    assert np.abs(got - want["counters"]).sum() <= 2
    # sharding invariance: two half-ranges give the same fates as one launch
    fate_full = res.fate.cpu().numpy()
    prop.reset()
    a = prop.propagate_philox(eng.make_source(vdist, xdist), 5, 1000, n // 2, want_fate=True).fate.cpu().numpy()
    b = prop.propagate_philox(eng.make_source(vdist, xdist), 5, 1000 + n // 2, n - n // 2, want_fate=True).fate.cpu().numpy()
    np.testing.assert_array_equal(np.concatenate([a, b]), fate_full)
    np.testing.assert_array_equal(prop.counters.cpu().numpy(), got)


def test_host_buffer_entry(torch_cuda, cuda_lib):
    """cmt_run_host_ic: plain host pointers in, fates/final rows/counters out."""
    import ctypes as C
    from trajectories import _engine as eng

    bl = lens_beamline(lens_table())
    n = (1 << 21) + 12345                             # more than one staging chunk, ragged tail
    ic = standard_ics(n, 31)
    prop = eng.Propagator(bl.elements, 0)
    fate = np.empty(n, dtype=np.uint8)
    fin = np.empty((10, n))
    counters = np.zeros(len(prop.flat.fate_names), dtype=np.int64)
    work = np.zeros(8, dtype=np.int64)
    rc = cuda_lib.cmt_run_host_ic(prop.dev.handle, n, ic.ctypes.data, fate.ctypes.data, fin.ctypes.data,
                                  counters.ctypes.data, work.ctypes.data)
    assert rc == 0, cuda_lib.cmt_last_error()
    want = oracle.propagate(bl.elements, ic)
    np.testing.assert_array_equal(fate, want["fate"])
    np.testing.assert_array_equal(counters, want["counters"])
    np.testing.assert_array_equal(work[:3], want["work"])
    assert relerr(fin, want["fin"]) < TIGHT


@pytest.mark.parametrize("case", ["", "early_", "spa_"])
def test_run_simulation_golden(torch_cuda, golden_dir, case):
    """TrajectorySimulator.run_simulation on replayed draws == the reference's own run (trajectory_simulator.py:34-103):
    the lens beamline saving Detected + Inside lens, the lens beamline saving EARLY fates (second aperture, lens
    entrance, and a name that matches nothing), and the SPA geometry with its Gaussian source."""
    from trajectories.distributions import Distribution
    from trajectories.trajectory_simulator import TrajectorySimulator

    g = np.load(golden_dir / "run_simulation.npz")

    class Replay(Distribution):
        def __init__(self, data):
            self.data, self.pos = data, 0

        def draw(self, n):
            out = self.data[:, self.pos:self.pos + n]
            self.pos += n
            return out

        def save_to_hdf(self, *a, **k):
            pass

    full = g
    g = {k[len(case):]: full[k] for k in full.files if k.startswith(case)} if case else full
    bl = spa_beamline() if case == "spa_" else lens_beamline((full["table_r"], full["table_a"]))
    sim = TrajectorySimulator()
    sim.run_simulation(bl, "golden", vdist=Replay(g["ic"][3:6]), xdist=Replay(g["ic"][0:3]),
                       N_traj=int(g["N_traj"]), apertures_of_interest=list(g["aoi"]), n_jobs=int(full["n_jobs"]))
    want = dict(zip(g["counter_keys"].tolist(), g["counter_vals"].tolist()))
    assert sim.counter.counter_dict == want
    assert sim.counter.calculate_efficiency() == float(g["efficiency"])
    saved = sim.result.molecules
    assert len(saved) == len(g["saved_fate"])
    assert [m.aperture_hit for m in saved] == g["saved_fate"].tolist()
    assert [m.alive for m in saved] == g["saved_alive"].tolist()
    assert [m.trajectory.x.shape[0] for m in saved] == g["saved_n_rows"].tolist()
    np.testing.assert_array_equal(np.array([m.trajectory.x[0] for m in saved]).T, g["saved_x0"])
    last = np.array([np.concatenate([m.trajectory.x[-1], m.trajectory.v[-1], m.trajectory.a[-1], [m.trajectory.t[-1]]])
                     for m in saved])
    assert relerr(last, g["saved_last"]) < TIGHT
    assert sim.results["golden"].counter is sim.counter
    m = saved[0]
    assert m.trajectory.x.shape[1] == 3 and m.trajectory.v.shape == m.trajectory.a.shape == m.trajectory.x.shape
    assert m.trajectory.t.shape == (m.trajectory.x.shape[0],) and m.trajectory.n == m.trajectory.x.shape[0]


def test_run_simulation_philox(torch_cuda):
    from trajectories.trajectory_simulator import TrajectorySimulator

    bl = lens_beamline(lens_table())
    sim = TrajectorySimulator(seed=3)
    sim.run_simulation(bl, "r", N_traj=1_000_050, apertures_of_interest=["Detected"], n_jobs=10)
    c = sim.counter.counter_dict
    assert sum(c.values()) == 1_000_000                       # remainder dropped like the reference
    assert abs(c["4K shield"] / 1e6 - 0.490) < 0.005 and abs(c["40K shield"] / 1e6 - 0.292) < 0.005
    assert abs(c["Lens entrance"] / 1e6 - 0.212) < 0.005
    assert len(sim.result.molecules) == c.get("Detected", 0)
    for m in sim.result.molecules:
        assert m.alive and m.aperture_hit == "Detected" and m.trajectory.x.shape == (613, 3)
    sim2 = TrajectorySimulator(seed=3)
    sim2.run_simulation(bl, "r", N_traj=1_000_050, apertures_of_interest=[], n_jobs=10)
    assert sim2.counter.counter_dict == c                      # same seed, same run


def test_plugin_single_molecule(torch_cuda, golden_dir):
    """Beamline.propagate_through(molecule) and element.propagate_through(molecule)."""
    from trajectories.molecule import Molecule

    g = np.load(golden_dir / "lens_biased.npz")
    bl = lens_beamline((g["table_r"], g["table_a"]))
    names = list(g["lens_fate_names"])
    off = g["lens_row_off"]
    for k in (0, 5, len(g["lens_row_idx"]) - 1):
        i = g["lens_row_idx"][k]
        want = g["lens_rows"][off[k]:off[k + 1]]
        m = Molecule()
        m.init_trajectory(bl, g["ic"][0:3, i], g["ic"][3:6, i])
        bl.propagate_through(m)
        assert m.aperture_hit == names[g["lens_fate"][i]] and m.alive == bool(g["lens_alive"][i])
        got = np.concatenate([m.trajectory.x, m.trajectory.v, m.trajectory.a, m.trajectory.t[:, None]], axis=1)
        assert got.shape == want.shape and relerr(got, want) < TIGHT
        # element by element, as Beamline.propagate_through does in the reference
        m2 = Molecule()
        m2.init_trajectory(bl, g["ic"][0:3, i], g["ic"][3:6, i])
        for e in bl.elements:
            e.propagate_through(m2)
            if not m2.alive:
                break
        m2.trajectory.drop_nans()
        got2 = np.concatenate([m2.trajectory.x, m2.trajectory.v, m2.trajectory.a, m2.trajectory.t[:, None]], axis=1)
        np.testing.assert_array_equal(got2, got)


def test_propagate_inside_lens(torch_cuda, golden_dir):
    """ElectrostaticLens.propagate_inside_lens (electrostatic_lens.py:79-118) used the way the reference's
    propagate_through uses it: row at z0 and entrance test by hand, the integration by the method, exit row by hand."""
    from trajectories.molecule import Molecule

    g = np.load(golden_dir / "lens_biased.npz")
    bl = lens_beamline((g["table_r"], g["table_a"]))
    lens = bl.find_element("ES lens")
    names = list(g["lens_fate_names"])
    off = g["lens_row_off"]
    through = 0
    for k in range(len(g["lens_row_idx"])):
        i = g["lens_row_idx"][k]
        want = g["lens_rows"][off[k]:off[k + 1]]
        m = Molecule()
        m.init_trajectory(bl, g["ic"][0:3, i], g["ic"][3:6, i])
        for e in bl.elements:
            if e is not lens:
                e.propagate_through(m)
            else:
                m.update_trajectory((lens.z0 - m.x()[2]) / m.v()[2])
                if np.sqrt(np.sum(m.x()[:2] ** 2)) > lens.d / 2:
                    m.set_dead()
                    m.set_aperture_hit("Lens entrance")
                else:
                    before = m.trajectory.n
                    lens.propagate_inside_lens(m)
                    through += 1
                    if m.alive:
                        assert m.trajectory.n - before == int(np.rint(lens.L / lens.dz))
                        m.update_trajectory((lens.z1 - m.x()[2]) / m.v()[2])
            if not m.alive:
                break
        if m.alive:
            m.set_aperture_hit("Detected")
        m.trajectory.drop_nans()
        assert m.aperture_hit == names[g["lens_fate"][i]] and m.alive == bool(g["lens_alive"][i])
        got = np.concatenate([m.trajectory.x, m.trajectory.v, m.trajectory.a, m.trajectory.t[:, None]], axis=1)
        assert got.shape == want.shape and relerr(got, want) < TIGHT
    assert through >= 2


def test_arithmetic_selftest(torch_cuda, cuda_lib):
    """The shared-reciprocal division and inline sqrt equal __ddiv_rn / __dsqrt_rn bit for bit."""
    import ctypes as C

    for mode, n in ((0, 100_000_000_000), (1, 100_000_000_000), (2, 100_000_000_000), (3, 100_000_000_000)):
        out = (C.c_int64 * 5)()
        assert cuda_lib.cmt_selftest(0, n, 0xC0FFEE + mode, mode, out) == 0, cuda_lib.cmt_last_error()
        took_div, bad_div, took_sqrt, bad_sqrt, bad_cached = list(out)
        assert bad_div == 0 and bad_sqrt == 0 and bad_cached == 0, (mode, list(out))
        if mode == 0:
            assert took_div == n and took_sqrt == n        # the short sequences cover the working range
        elif mode == 1:
            assert took_div == n                           # w/6 in two operations: every operand StepCheck admits
        elif mode == 2:
            assert 0.5 * n < took_div < n and took_sqrt > 0.9 * n   # extremes fall back
        else:
            assert took_div > 0.3 * n and took_sqrt == took_div     # a / sqrt(s) with the root's own reciprocal


def test_fast_math_equals_reference_math(torch_cuda, cuda_lib):
    """Kernels with shared reciprocals / FMA-carried power-of-two scalings give the same
    bits as the same kernels forced onto the plain-intrinsic path."""
    from trajectories import _engine as eng

    torch = torch_cuda
    bl = lens_beamline(lens_table())
    ic = torch.from_numpy(standard_ics(300000, 77, 3.0)).cuda()
    flat = eng.flatten(bl.elements)
    outs = []
    for flags in (0, 1):                                    # 1: plain-intrinsic arithmetic
        old = cuda_lib.cmt_debug_flags(flags)
        try:
            prop = eng.Propagator(flat, 0)
            prop.dev = eng.DeviceBeamline(flat, 0)          # bypass the handle cache: flags are read at creation
            prop.reset()
            res = prop.propagate_ic(ic, want_fate=True, want_final=True)
            rows, offs, fate = prop.trajectories(ic[:, :2000].contiguous())
            torch.cuda.synchronize()
            outs.append((res.fate.cpu().numpy(), res.final.cpu().numpy(), res.work.cpu().numpy(), rows, offs))
        finally:
            cuda_lib.cmt_debug_flags(old)
    np.testing.assert_array_equal(outs[0][0], outs[1][0])
    np.testing.assert_array_equal(outs[0][1].view(np.int64), outs[1][1].view(np.int64))
    np.testing.assert_array_equal(outs[0][2][:4], outs[1][2][:4])
    # RK steps on the plain-intrinsic path: a few per 1e7 (a radius within 2^-20 of its table interval's end fails
    # the one-comparison interval test and is redone there) / all
    assert outs[0][2][4] <= 1e-5 * outs[0][2][1] and outs[1][2][4] == outs[1][2][1]
    np.testing.assert_array_equal(outs[0][4], outs[1][4])
    np.testing.assert_array_equal(outs[0][3].view(np.int64), outs[1][3].view(np.int64))
    assert outs[0][2][1] > 50_000_000                      # tens of millions of RK steps compared


def test_run_simulation_spa_saved_trajectories(torch_cuda):
    """BASELINE.json configs[3]: SPA beamline, Gaussian position source, detected molecules saved."""
    from trajectories.distributions import GaussianPositionDistribution
    from trajectories.trajectory_simulator import TrajectorySimulator

    bl = spa_beamline()
    sim = TrajectorySimulator(seed=11, chunk=1 << 20)           # several chunks
    sim.run_simulation(bl, "spa", N_traj=int(3e6), apertures_of_interest=["Detected"], n_jobs=9,
                       xdist=GaussianPositionDistribution())
    c = sim.counter.counter_dict
    n_run = 900 * int(3e6 / 900)
    assert sum(c.values()) == n_run
    eff = sim.counter.calculate_efficiency()
    assert abs(eff - 3.1e-4) < 6e-5                              # SURVEY.md 3.4 [probe]: ~3.1e-4
    mols = sim.result.molecules
    assert len(mols) == c["Detected"]
    for m in mols[:50]:
        assert m.trajectory.x.shape == (19, 3) and m.trajectory.t.shape == (19,)   # 1 + 2 x 9 rows
        assert m.alive and m.aperture_hit == "Detected"
        assert np.all(np.diff(m.trajectory.t) >= 0) and np.all(np.diff(m.trajectory.x[:, 2]) >= 0)
        assert np.all(m.trajectory.a[:, 1] == -9.80665) and np.all(m.trajectory.a[:, [0, 2]] == 0)
    # the saved trajectories are exactly what the oracle computes from their first rows
    ic = np.array([np.concatenate([m.trajectory.x[0], m.trajectory.v[0]]) for m in mols[:200]]).T
    want = oracle.propagate(bl.elements, ic, want_rows=True)
    for k, m in enumerate(mols[:200]):
        got = np.concatenate([m.trajectory.x, m.trajectory.v, m.trajectory.a, m.trajectory.t[:, None]], axis=1)
        assert relerr(got, want["rows"][k, :19]) < TIGHT
    # same seed without saving: identical Counter (the saved-index path does not disturb the run)
    sim2 = TrajectorySimulator(seed=11)
    sim2.run_simulation(bl, "spa", N_traj=int(3e6), n_jobs=9, xdist=GaussianPositionDistribution())
    assert sim2.counter.counter_dict == c


def test_state_and_voltage_sweep(torch_cuda):
    """BASELINE.json configs[2] in miniature: a loop over states and voltages like
    examples/lens_simulation_different_states.py:137-152 (state swapped in place, a_interp reset)."""
    from trajectories.stark_potential import UncoupledBasisState
    from trajectories.trajectory_simulator import TrajectorySimulator

    bl = lens_beamline()                                           # table built lazily from lens.state / lens.V
    sim = TrajectorySimulator(seed=5)
    eff = {}
    for (J, mJ) in [(0, 0), (1, 0), (2, 0), (2, 2)]:
        for V in (20e3, 30e3):
            lens = bl.find_element("ES lens")
            lens.state = 1 * UncoupledBasisState(J=J, mJ=mJ, I1=1 / 2, m1=1 / 2, I2=1 / 2, m2=1 / 2, Omega=0,
                                                 P=(-1) ** J, electronic_state="X")
            lens.V = V
            lens.a_interp = None
            sim.run_simulation(bl, f"J={J},mJ={mJ},V={V}", N_traj=int(2e6), apertures_of_interest=["Detected"], n_jobs=10)
            eff[(J, mJ, V)] = sim.counter.calculate_efficiency()
            assert sum(sim.counter.counter_dict.values()) == int(2e6)
    assert len(sim.results) == 8
    # J=0 is high-field seeking (defocused): far fewer detected molecules than the focused J=2, mJ=0
    assert eff[(0, 0, 30e3)] < 0.5 * eff[(2, 0, 30e3)]
    assert eff[(2, 0, 30e3)] > 1e-4


def _same_bits_or_close(a, b, tol=TIGHT):
    """NaN-aware comparison: identical NaN/Inf pattern, finite values within tol."""
    a, b = np.asarray(a), np.asarray(b)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.array_equal(np.isinf(a), np.isinf(b)) and np.array_equal(np.sign(a[np.isinf(a)]), np.sign(b[np.isinf(b)]))
    f = np.isfinite(a)
    if f.any():
        assert relerr(a[f], b[f]) < tol


def test_unusual_beamlines(torch_cuda):
    """Two lenses in series, a lens first, a lens with a coarse user table on an uneven grid
    (takes the plain-intrinsic path), overlapping elements, an empty beamline."""
    from trajectories.beamline import Beamline
    from trajectories.beamline_elements.apertures import CircularAperture, FieldPlates, RectangularAperture
    from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens, make_interpolator

    table = lens_table()
    ic = standard_ics(40000, 5, 2.5)

    def lens(name, z0, L, tab=table, **kw):
        return ElectrostaticLens(name=name, z0=z0, L=L, a_interp=make_interpolator(*tab), **kw)

    uneven_r = np.array([0.0, 0.001, 0.0035, 0.004, 0.011, 0.0199, 0.0225])
    uneven = (uneven_r, -2.0e5 * uneven_r ** 1.1)
    cases = {
        "two lenses": [CircularAperture(name="in", z0=0.05, L=0.01, d=0.03), lens("L1", 0.4, 0.3),
                       FieldPlates(name="fp", z0=0.8, L=0.2, w=0.03), lens("L2", 1.2, 0.25, dz=2e-3),
                       RectangularAperture(name="out", z0=2.0, L=0.01, w=0.02, h=0.02)],
        "lens first": [lens("L", 0.1, 0.5), CircularAperture(name="c", z0=1.0, L=0.1, d=0.02)],
        "uneven table": [CircularAperture(name="in", z0=0.05, L=0.01, d=0.03), lens("L", 0.4, 0.3, tab=uneven)],
        "overlap": [CircularAperture(name="a", z0=0.10, L=0.30, d=0.03), CircularAperture(name="b", z0=0.20, L=0.05, d=0.02),
                    RectangularAperture(name="r", z0=0.21, L=0.5, w=0.05, h=0.01)],
        "lens only, one step": [lens("L", 0.3, 1e-3)],
        # 2500 points = 80 kB of shared memory: needs the opt-in dynamic shared-memory limit
        "fine table": [CircularAperture(name="in", z0=0.05, L=0.01, d=0.03),
                       lens("L", 0.4, 0.3, tab=(np.linspace(0, 0.0225, 2500), -3.0e4 * np.linspace(0, 0.0225, 2500) ** 0.9))],
    }
    for name, elems in cases.items():
        bl = Beamline(elems)
        want = oracle.propagate(bl.elements, ic)
        got = gpu_propagate(torch_cuda, bl, ic)
        np.testing.assert_array_equal(got["fate"], want["fate"], err_msg=name)
        np.testing.assert_array_equal(got["counters"], want["counters"], err_msg=name)
        np.testing.assert_array_equal(got["work"][:3], want["work"], err_msg=name)
        assert relerr(got["fin"], want["fin"]) < TIGHT, name
        if name == "uneven table":
            assert got["work"][4] > 0.5 * got["work"][1]     # uneven grids use the plain-intrinsic step
        # full trajectories of a sample through the same beamline
        rows, offs, fate = got["prop"].trajectories(torch_cuda.from_numpy(np.ascontiguousarray(ic[:, :300])).cuda())
        w2 = oracle.propagate(bl.elements, ic[:, :300], want_rows=True)
        np.testing.assert_array_equal(np.diff(offs), w2["n_rows"], err_msg=name)
        for k in range(300):
            assert relerr(rows[offs[k]:offs[k + 1]], w2["rows"][k, : w2["n_rows"][k]]) < TIGHT, name
    # an empty beamline detects everything and leaves the initial row untouched
    got = gpu_propagate(torch_cuda, Beamline([]), ic[:, :1000])
    assert (got["fate"] == 0).all() and got["counters"].tolist() == [1000]
    np.testing.assert_array_equal(got["fin"][0:6], ic[:, :1000])


@pytest.mark.parametrize("n_steps", [1, 149, 150, 151, 301, 2100, 2101, 3000])
def test_lens_segment_boundaries(torch_cuda, n_steps):
    """The lens integrator runs as launches of 150 RK steps (at most 14 launches: longer lenses get longer
    segments); step counts on, next to and far beyond the segment boundaries, with elements behind the lens."""
    from trajectories.beamline import Beamline
    from trajectories.beamline_elements.apertures import CircularAperture, FieldPlates
    from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens, make_interpolator

    L = 0.6
    lens = ElectrostaticLens(name="L", z0=0.4, L=L, dz=L / n_steps, a_interp=make_interpolator(*lens_table()))
    assert lens.N_steps() == n_steps + 1                      # its rows: one per RK step and the exit row
    bl = Beamline([CircularAperture(name="in", z0=0.05, L=0.01, d=0.03), lens,
                   FieldPlates(name="fp", z0=1.1, L=0.2, w=0.03), CircularAperture(name="out", z0=1.5, L=0.01, d=0.02)])
    ic = standard_ics(6000, 40 + n_steps % 7, 2.5)
    want = oracle.propagate(bl.elements, ic)
    got = gpu_propagate(torch_cuda, bl, ic)
    assert want["work"][1] > 1000 * min(n_steps, 50)          # the lens is actually exercised
    np.testing.assert_array_equal(got["fate"], want["fate"])
    np.testing.assert_array_equal(got["counters"], want["counters"])
    np.testing.assert_array_equal(got["work"][:3], want["work"])
    assert relerr(got["fin"], want["fin"]) < TIGHT


def test_table_out_of_range_is_counted_and_raised(torch_cuda):
    """A force evaluation beyond the a_interp table: the reference's interp1d raises ValueError
    (electrostatic_lens.py:209,217); here it is counted on the device and raised after the run."""
    from trajectories.beamline import Beamline
    from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens, make_interpolator
    from trajectories.distributions import Distribution
    from trajectories.trajectory_simulator import TrajectorySimulator

    r = np.linspace(0, 0.006, 30)                      # table ends well inside the 22 mm bore
    bl = Beamline([ElectrostaticLens(name="L", z0=0.1, L=0.3, a_interp=make_interpolator(r, -1e4 * r))])
    ic = standard_ics(5000, 3, 5.0)
    want = oracle.propagate(bl.elements, ic)
    got = gpu_propagate(torch_cuda, bl, ic)
    assert want["work"][2] > 0
    np.testing.assert_array_equal(got["work"][:3], want["work"])       # same count of out-of-range evaluations
    np.testing.assert_array_equal(got["fate"], want["fate"])
    assert relerr(got["fin"], want["fin"]) < TIGHT

    class Fixed(Distribution):
        def __init__(self, data):
            self.data = data

        def draw(self, n):
            return self.data[:, :n]

        def save_to_hdf(self, *a, **k):
            pass

    with pytest.raises(ValueError, match="interpolation range"):
        TrajectorySimulator().run_simulation(bl, "r", vdist=Fixed(ic[3:6]), xdist=Fixed(ic[0:3]), N_traj=4000)


def test_hostile_initial_conditions(torch_cuda):
    """vz = 0, negative vz, NaN/Inf components, huge and tiny values: the GPU follows the same IEEE
    comparisons as the reference (e.g. `rho > d/2` is False for NaN), so fates still match the oracle."""
    bl = lens_beamline(lens_table())
    base = standard_ics(64, 9, 3.0)
    ic = np.repeat(base, 12, axis=1)
    n = ic.shape[1]
    k = np.arange(n) % 12
    ic[5, k == 1] = 0.0                      # vz = 0 -> dt = +-inf
    ic[5, k == 2] *= -1                      # flying backwards: negative dt
    ic[0, k == 3] = np.nan
    ic[4, k == 4] = np.inf
    ic[5, k == 5] = 1e-300                   # dt overflows
    ic[5, k == 6] = 1e300                    # dt underflows towards 0
    ic[0, k == 7] = 1e200
    ic[2, k == 8] = lens_beamline(lens_table()).elements[0].z0      # starts exactly on the first plane: dt == 0
    ic[3, k == 9] = -0.0
    ic[1, k == 10] = 5e-324                  # subnormal position
    ic[5, k == 11] = np.nan
    want = oracle.propagate(bl.elements, ic)
    got = gpu_propagate(torch_cuda, bl, ic)
    np.testing.assert_array_equal(got["fate"], want["fate"])
    np.testing.assert_array_equal(got["counters"], want["counters"])
    _same_bits_or_close(got["fin"], want["fin"])
    ap = apertures_beamline()
    want = oracle.propagate(ap.elements, ic)
    got = gpu_propagate(torch_cuda, ap, ic)
    np.testing.assert_array_equal(got["fate"], want["fate"])
    _same_bits_or_close(got["fin"], want["fin"])


def test_full_size_properties(torch_cuda):
    """BASELINE.json configs[1] at full size (1e7 molecules): size-independent checks."""
    from trajectories import _engine as eng
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution

    torch = torch_cuda
    bl = lens_beamline(lens_table())
    n = 10_000_000
    prop = eng.Propagator(bl.elements, 0)
    src = eng.make_source(CeNTREXVelocityDistribution(), CeNTREXPositionDistribution())
    names = prop.flat.fate_names
    # (1) the source run and the replay of the same samples agree molecule by molecule
    prop.reset()
    a = prop.propagate_philox(src, 77, 5_000_000_000, n, want_fate=True)
    fate_a, cnt_a = a.fate.clone(), a.counters.clone()
    ic = prop.draw(src, 77, 5_000_000_000, n)
    prop.reset()
    b = prop.propagate_ic(ic, first_index=5_000_000_000, want_fate=True, want_final=True,
                          save_mask=1 << names.index("Detected"))
    assert torch.equal(fate_a, b.fate) and torch.equal(cnt_a, b.counters)
    # (2) the Counter is the histogram of the fate bytes and sums to N
    hist = torch.bincount(b.fate.long(), minlength=len(names))
    assert torch.equal(hist, b.counters) and int(b.counters.sum()) == n
    # (3) the saved-index list is exactly the detected molecules, sorted, in global numbering
    det = torch.nonzero(b.fate == names.index("Detected")).flatten() + 5_000_000_000
    assert torch.equal(det, b.saved_index)
    # (4) geometry of the final rows: every fate ends on the plane (or wall) that defines it
    fin = b.final
    els = {e.name: e for e in bl.elements}
    sel = b.fate == names.index("Detected")
    assert torch.allclose(fin[2, sel], torch.full_like(fin[2, sel], els["DR aperture"].z1), rtol=0, atol=1e-12)
    assert bool((fin[0, sel].abs() < 0.009).all()) and bool((fin[1, sel].abs() < 0.015).all())
    sel = b.fate == names.index("Lens entrance")
    assert torch.allclose(fin[2, sel], torch.full_like(fin[2, sel], els["ES lens"].z0), rtol=0, atol=1e-12)
    assert bool((torch.hypot(fin[0, sel], fin[1, sel]) > els["ES lens"].d / 2).all())
    sel = b.fate == names.index("Inside lens")
    assert bool((torch.hypot(fin[0, sel], fin[1, sel]) > els["ES lens"].d / 2).all())
    assert bool((fin[2, sel] > els["ES lens"].z0).all()) and bool((fin[2, sel] < els["ES lens"].z1 + 1e-9).all())
    assert bool((fin[6, sel] != 0).any())                       # the stored a is the lens force l1, not (0,-g,0)
    sel = b.fate == names.index("Field plates")
    on_wall = (fin[0, sel].abs() - 0.01).abs() < 1e-12          # stopped where it crossed a plate ...
    at_entry = (fin[2, sel] - els["Field plates"].z0).abs() < 1e-12   # ... or was outside at z0
    assert bool((on_wall | at_entry).all())
    # (5) time and energy bookkeeping: t > 0, vz untouched, vy = vy0 - g t outside the lens
    never_lens = b.fate < names.index("Inside lens")
    assert bool((fin[9] > 0).all()) and torch.equal(fin[5], ic[5])
    vy_pred = ic[4, never_lens] - 9.80665 * fin[9, never_lens]
    assert float((fin[4, never_lens] - vy_pred).abs().max()) < 1e-12
    # (6) work counters: rows + steps account for every trajectory row
    w = b.work.cpu().numpy()
    assert w[3] == int(b.counters[names.index("Inside lens"):].sum())      # lens entries = everything after the entrance
    assert w[2] == 0 and w[4] <= 1e-5 * w[1]       # no table excursions; a few RK steps per 1e7 redone on the plain path


def test_hit_fractions_match_the_reference(torch_cuda, golden_dir):
    """Per-element hit fractions of an independent Philox run agree with the fractions of the
    reference's own run (golden std sets, 12 000 molecules) within Monte Carlo error."""
    from trajectories.trajectory_simulator import TrajectorySimulator

    ref_counts, names = None, None
    for seed in (0, 1, 2):
        g = np.load(golden_dir / f"std_seed{seed}.npz")
        names = list(g["lens_fate_names"])
        c = np.bincount(g["lens_fate"], minlength=len(names))
        ref_counts = c if ref_counts is None else ref_counts + c
    n_ref = ref_counts.sum()
    bl = lens_beamline((g["table_r"], g["table_a"]))
    sim = TrajectorySimulator(seed=123)
    sim.run_simulation(bl, "r", N_traj=int(2e7), n_jobs=10)
    n = sum(sim.counter.counter_dict.values())
    for k, name in enumerate(names):
        p = sim.counter.counter_dict.get(name, 0) / n             # 2e7 molecules: essentially the true fraction
        sigma = np.sqrt(max(p * (1 - p), 1e-12) / n_ref)
        assert abs(ref_counts[k] / n_ref - p) < 4.5 * sigma + 1.5 / n_ref, (name, ref_counts[k] / n_ref, p)


def test_saving_an_early_fate_is_compact(torch_cuda):
    """Saving molecules that die at the first aperture (2-3 rows each, half of all molecules) must not
    allocate 613 rows per molecule."""
    from trajectories.trajectory_simulator import TrajectorySimulator

    bl = lens_beamline(lens_table())
    sim = TrajectorySimulator(seed=2)
    sim.run_simulation(bl, "r", N_traj=400_000, apertures_of_interest=["4K shield", "Detected"], n_jobs=10)
    c = sim.counter.counter_dict
    mols = sim.result.molecules
    assert len(mols) == c["4K shield"] + c.get("Detected", 0)
    n_rows = np.array([m.trajectory.x.shape[0] for m in mols])
    hit = np.array([m.aperture_hit == "4K shield" for m in mols])
    assert set(n_rows[hit]) <= {2, 3} and (n_rows[~hit] == 613).all()
    for m in mols[:100]:
        assert not m.alive and m.trajectory.t.shape == (m.trajectory.x.shape[0],)
        r = np.hypot(*m.trajectory.x[-1, :2])
        assert r > 0.0127                                             # it ended outside the 1-inch aperture


def test_graph_replay_and_stream_slots(torch_cuda):
    """A step captured into a CUDA graph, and steps issued on the private stream slots, give
    exactly what the plain call on the current stream gives."""
    from trajectories import _engine as eng

    torch = torch_cuda
    bl = lens_beamline(lens_table())
    ic = torch.from_numpy(standard_ics(500_000, 41, 6.0)).cuda()
    prop = eng.Propagator(bl.elements, 0)
    prop.reset()
    ref = prop.propagate_ic(ic, want_fate=True)
    torch.cuda.synchronize()
    fate_ref, cnt_ref, work_ref = ref.fate.clone(), ref.counters.clone(), ref.work.clone()
    # stream slots
    prop.reset()
    outs = [prop.propagate_ic(ic, want_fate=True, slot=k) for k in range(6)]
    prop.join()
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o.fate, fate_ref)
    assert torch.equal(prop.counters, 6 * cnt_ref) and torch.equal(prop.work, 6 * work_ref)
    # graph replay
    prop.reset()
    steps = [prop.capture_ic(ic, want_fate=True, slot=s) for s in range(prop.n_slots)]
    prop.reset()                                   # capture itself does not run the kernels, but be explicit
    for k in range(7):
        steps[k % len(steps)].replay()
    prop.join()
    torch.cuda.synchronize()
    for s in steps:
        assert torch.equal(s.fate, fate_ref)
    assert torch.equal(prop.counters, 7 * cnt_ref) and torch.equal(prop.work, 7 * work_ref)


def test_integration_stub_from_the_docs(torch_cuda):
    """The ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would add)
    is executable as written and reproduces the oracle."""
    import re
    from pathlib import Path

    from trajectories import _native

    root = Path(__file__).resolve().parent.parent
    text = (root / "INTEGRATION.md").read_text()
    block = re.search(r"```python\n# trajectories/_cmt.py.*?\n(.*?)```", text, re.S).group(1)
    block = block.replace('C.CDLL("libcmt_b200.so")', f'C.CDLL("{_native.LIB_PATH}")')
    ns = {}
    exec(compile(block, "INTEGRATION.md:_cmt.py", "exec"), ns)
    bl = lens_beamline(lens_table())
    ic = standard_ics(30000, 8, 4.0)
    fate, counts = ns["propagate"](bl, ic[0:3], ic[3:6])
    want = oracle.propagate(bl.elements, ic)
    np.testing.assert_array_equal(fate, want["fate"])
    assert counts == {nm: int(c) for nm, c in zip(want["fate_names"], want["counters"]) if c}


LOOSE = 1e-9        # north_star: ~1e-9 on ballistic segments, ~1e-6 through the lens


def test_contracted_math_mode(torch_cuda, golden_dir):
    """math="contracted": the same algorithm with fused multiply-adds and reciprocal multiplications.
    Stated tolerance: final rows within 1e-9 relative of the reference (measured: ~1e-13); fates equal
    except for molecules within ~1e-12 (relative) of the edge that decides them."""
    from trajectories.trajectory_simulator import TrajectorySimulator

    # against the reference's own outputs
    for name in ("std_seed0", "lens_biased", "lens_biased_J1m1_20kV", "edges"):
        g = np.load(golden_dir / f"{name}.npz")
        got = gpu_propagate(torch_cuda, lens_beamline((g["table_r"], g["table_a"])), g["ic"], math="contracted")
        same = got["fate"] == g["lens_fate"]
        assert same.sum() >= len(same) - (2 if name == "edges" else 0), name     # "edges" sits on the boundaries
        assert relerr(got["fin"][:, same], g["lens_fin"][:, same]) < LOOSE
    g = np.load(golden_dir / "spa.npz")
    got = gpu_propagate(torch_cuda, spa_beamline(), g["ic"], math="contracted")
    np.testing.assert_array_equal(got["fate"], g["spa_fate"])
    assert relerr(got["fin"], g["spa_fin"]) < LOOSE
    # against the exact mode at a size where edge cases would show up
    bl = lens_beamline(lens_table())
    for n, seed, sigma in ((1_000_000, 61, 39.5), (300_000, 62, 3.0)):
        ic = standard_ics(n, seed, sigma)
        a = gpu_propagate(torch_cuda, bl, ic)
        b = gpu_propagate(torch_cuda, bl, ic, math="contracted")
        differ = np.nonzero(a["fate"] != b["fate"])[0]
        assert len(differ) <= 2, len(differ)
        same = a["fate"] == b["fate"]
        err = np.abs(b["fin"][:, same] - a["fin"][:, same]) / np.maximum(np.abs(a["fin"][:, same]), 1e-9)
        assert err.max() < LOOSE
        assert np.percentile(err, 99.9) < 1e-11                      # what is actually achieved
        np.testing.assert_array_equal(a["work"][2:4], b["work"][2:4])   # no table excursions, the same lens entries
        assert b["work"][4] == 0 and a["work"][4] <= 1e-5 * a["work"][1]  # exact mode redoes a few RK steps per 1e7 on the plain path
    # the public API in this mode: saved trajectories are re-propagated with the same arithmetic
    sim = TrajectorySimulator(seed=3, math="contracted")
    sim.run_simulation(bl, "r", N_traj=2_000_000, apertures_of_interest=["Detected", "Inside lens"], n_jobs=10)
    c = sim.counter.counter_dict
    assert len(sim.result.molecules) == c["Detected"] + c["Inside lens"]
    assert sum(m.aperture_hit == "Detected" for m in sim.result.molecules) == c["Detected"]
    assert all(m.trajectory.x.shape[0] == 613 for m in sim.result.molecules if m.alive)
    exact = TrajectorySimulator(seed=3)
    exact.run_simulation(bl, "r", N_traj=2_000_000, n_jobs=10)
    assert sum(abs(exact.counter.counter_dict[k] - c.get(k, 0)) for k in exact.counter.counter_dict) <= 4


def test_radius_threshold_is_exact(torch_cuda):
    """`sqrt(x^2+y^2) > d/2` is evaluated on the GPU as `x^2+y^2 > T` with a precomputed T: the two
    must agree for every double, in particular within a few ulp of the edge."""
    from trajectories.beamline import Beamline
    from trajectories.beamline_elements.apertures import CircularAperture

    rng = np.random.default_rng(99)
    hits = total = 0
    for trial in range(40):
        d = float(rng.uniform(1e-4, 0.3)) if trial % 4 else float(2.0 ** rng.integers(-12, -1))
        R = d / 2
        bl = Beamline([CircularAperture(name="c", z0=0.5, L=0.0, d=d)])
        n = 4000
        theta = rng.uniform(0, 2 * np.pi, n)
        rho = np.full(n, R)
        for _ in range(int(rng.integers(0, 4))):                 # a few ulp inside / outside
            rho = np.nextafter(rho, np.where(rng.random(n) < 0.5, 0.0, 1.0))
        ic = np.zeros((6, n))
        ic[0], ic[1] = rho * np.cos(theta), rho * np.sin(theta)
        ic[0, : n // 4], ic[1, : n // 4] = rho[: n // 4], 0.0    # on an axis: x^2 + 0 exactly
        ic[2], ic[5] = 0.5, 200.0                                # starts on the plane: dt = 0, position unchanged
        want = oracle.propagate(bl.elements, ic)
        got = gpu_propagate(torch_cuda, bl, ic)
        np.testing.assert_array_equal(got["fate"], want["fate"])
        hits += int(want["counters"][0])
        total += n
    assert 0.2 < hits / total < 0.8                              # both outcomes are well represented


def test_abi_from_plain_c(torch_cuda):
    """examples/abi_demo.c drives libcmt_b200.so from C (dlopen, no Python/torch in that process);
    its Counter equals the Python path's for the same beamline, table, source and seed."""
    import json
    import subprocess

    import __graft_entry__ as ge
    from trajectories import _native
    from trajectories.beamline_elements.electrostatic_lens import make_interpolator
    from trajectories.trajectory_simulator import TrajectorySimulator

    demo = ge.build_abi_demo()
    out = subprocess.run([str(demo), str(_native.LIB_PATH), "2000000", "7"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout)
    bl = lens_beamline()
    lens = bl.find_element("ES lens")
    r = np.array([1.01 * (lens.d / 2) * i / 221 for i in range(222)])
    lens.a_interp = make_interpolator(r, -2.0e4 * r)
    sim = TrajectorySimulator(seed=7)
    sim.run_simulation(bl, "r", N_traj=2_000_000, n_jobs=10)
    steps = int(sim.last_work[1])
    want = dict(sim.counter.counter_dict)
    assert {k: v for k, v in got.items() if v and k != "lens_rk_steps"} == want
    assert got["lens_rk_steps"] == steps


def test_run_size_arithmetic_of_the_reference(torch_cuda):
    """trajectory_simulator.py:48-49: N = int(N_traj / (100 n_jobs)) per loop, remainder dropped;
    N_traj below 100 n_jobs simulates nothing; unknown apertures of interest never match."""
    from trajectories.trajectory_simulator import TrajectorySimulator

    bl = apertures_beamline()
    sim = TrajectorySimulator(seed=1)
    sim.run_simulation(bl, "small", N_traj=999, n_jobs=10)              # N_loops = 1000 > N_traj
    assert sim.counter.counter_dict == {} and sim.result.molecules == [] and sim.counter.calculate_efficiency() == 0
    sim.run_simulation(bl, "odd", N_traj=12_345, n_jobs=3, apertures_of_interest=["no such element"])
    assert sum(sim.counter.counter_dict.values()) == 300 * 41          # int(12345 / 300) = 41 per loop
    assert sim.result.molecules == []
    assert set(sim.results) == {"small", "odd"} and sim.results["odd"].counter is sim.counter
    sim.run_simulation(bl, "float", N_traj=1e4, n_jobs=1)               # scripts pass floats (argparse type=float)
    assert sum(sim.counter.counter_dict.values()) == 10_000


def test_random_beamlines(torch_cuda):
    """Randomised geometry (profiles/fuzz_geometry.py: element types, sizes, positions with overlaps, lens tables and
    steps, a source scaled to the apertures): binary64 walk and all three forms of the FP32 filter against the oracle."""
    import importlib.util
    from pathlib import Path

    from trajectories import _engine as eng
    from trajectories import _native as nat

    spec = importlib.util.spec_from_file_location("fuzz_geometry", Path(__file__).resolve().parent.parent / "profiles" / "fuzz_geometry.py")
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    rng = np.random.default_rng(2026)
    lens_cases = 0
    for c in range(24):
        problems, info = fuzz.run_case(torch_cuda, oracle, eng, nat, rng, 8000)
        assert not problems, (c, problems, info)
        lens_cases += any(t == "ElectrostaticLens" for t, _, _ in info["elements"])
    assert lens_cases >= 5
