"""BASELINE.json configs[2] and configs[3] at their full sizes, through the public API, with the Counters checked
against the CPU oracle on the same seed and global-index range.

configs[2]  lens_simulation_different_states sweep: 8 rotational states x 5 lens voltages, 1e7 molecules each
            (examples/lens_simulation_different_states.py:72-152; the voltage axis is the build's)
configs[3]  examples/SPA beamline, Gaussian position source, 1e9 molecules, detected trajectories saved
            (examples/SPA/SPA_distributions.py:21-103)

The source is Philox indexed by the global molecule number; its integer stream is bit-identical between oracle and
device, its transcendental transforms differ by ulps (CUDA libm vs glibc), which can move a molecule across an
edge once in ~1e7: per-fate counts may differ by a few molecules, nothing else (tolerance stated in each test)."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle
from tests.beamlines import lens_beamline, spa_beamline

STATES = [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1)]        # first eight of the example's list
VOLTAGES = [20e3, 24e3, 27.6e3, 30e3, 34e3]
SLACK = 6          # molecules per fate and 1e7 that may sit on the other side of an edge (libm ulps)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


def counter_close(got: dict, names, want_counts, n, slack=SLACK):
    want = {nm: int(c) for nm, c in zip(names, want_counts) if c > 0}
    assert sum(got.values()) == sum(want.values()) == n
    for nm in set(got) | set(want):
        assert abs(got.get(nm, 0) - want.get(nm, 0)) <= slack, (nm, got, want)


def test_config2_state_and_voltage_sweep_full_size(torch_cuda):
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution
    from trajectories.stark_potential import UncoupledBasisState
    from trajectories.trajectory_simulator import TrajectorySimulator

    def state(J, mJ):
        return 1 * UncoupledBasisState(J=J, mJ=mJ, I1=1 / 2, m1=1 / 2, I2=1 / 2, m2=-1 / 2, Omega=0, P=(-1) ** J,
                                       electronic_state="X")

    n, seed = 10_000_000, 7
    bl = lens_beamline()
    sim = TrajectorySimulator(seed=seed)
    sim.run_sweep(bl, [state(1, 0)], [25e3], N_traj=n, n_jobs=10)            # warm-up: library, streams, workspaces
    torch_cuda.cuda.synchronize()
    seconds = []
    for _ in range(2):                  # best of two: the assertion is about the path, not about a cold allocator
        t0 = time.perf_counter()
        res = sim.run_sweep(bl, [state(J, mJ) for J, mJ in STATES], VOLTAGES, N_traj=n, n_jobs=10)
        seconds.append(time.perf_counter() - t0)
    print("configs[2] run_sweep wall times:", [round(s, 3) for s in seconds])
    seconds = min(seconds)
    assert len(res) == 40 and set(res) == {(J, mJ, V) for J, mJ in STATES for V in VOLTAGES}
    assert len(sim.results) >= 40 and "J = 2, mJ = 0, V = 27600" in sim.results
    print(f"configs[2]: 4e8 molecules, 40 sweep points in {seconds:.3f} s = {4e8 / seconds:.3g} molecules/s through run_sweep")
    assert seconds < 1.0                                                      # 1.65 s as 40 run_simulation calls in round 1

    # every point against the oracle: same seed, same global index range [0, 1e7)
    src = oracle.make_source(CeNTREXVelocityDistribution(), CeNTREXPositionDistribution())
    front = ["4K shield", "40K shield", "BB exit", "Lens entrance"]
    first = None
    for key, r in res.items():
        want = oracle.run(r.beamline.elements, src, seed, 0, n, n_threads=oracle.host_cores())
        counter_close(r.counter.counter_dict, want["fate_names"], want["counters"], n)
        # the walk up to the lens entrance does not depend on the lens: identical counts at every point
        head = [r.counter.counter_dict.get(k, 0) for k in front]
        first = first or head
        assert head == first
    # focusing physics of the rigid-rotor curve: J=2, mJ=0 (low-field seeker) beats J=0 at every voltage
    for V in VOLTAGES:
        assert res[(2, 0, V)].counter.calculate_efficiency() > 2 * res[(0, 0, V)].counter.calculate_efficiency()

    # the batched call equals the example's own loop: state swapped in place, table reset, run_simulation
    lens = bl.find_element("ES lens")
    for J, mJ, V in [(0, 0, 20e3), (2, 0, 27.6e3), (3, 1, 34e3)]:
        lens.state, lens.V, lens.a_interp = state(J, mJ), V, None
        one = TrajectorySimulator(seed=seed)
        one.run_simulation(bl, "loop", N_traj=n, n_jobs=10)
        assert one.counter.counter_dict == res[(J, mJ, V)].counter.counter_dict

    # with trajectories: the same saved molecules as the loop gives
    lens.state, lens.V, lens.a_interp = state(2, 0), 27.6e3, None
    one = TrajectorySimulator(seed=seed)
    one.run_simulation(bl, "loop", N_traj=2_000_000, apertures_of_interest=["Detected"], n_jobs=10)
    swept = TrajectorySimulator(seed=seed).run_sweep(bl, [state(2, 0), state(1, 1)], [27.6e3], N_traj=2_000_000,
                                                      apertures_of_interest=["Detected"], n_jobs=10)
    a, b = one.result.molecules, swept[(2, 0, 27.6e3)].molecules
    assert len(a) == len(b) > 100
    for ma, mb in zip(a, b):
        np.testing.assert_array_equal(ma.trajectory.x, mb.trajectory.x)
    # run_sweep leaves the caller's lens as it found it (here: the state, voltage and table of the loop run above)
    assert lens.V == 27.6e3 and lens.state == state(2, 0) and lens.a_interp is not None
    np.testing.assert_array_equal(lens.a_interp.y, swept[(2, 0, 27.6e3)].beamline.find_element("ES lens").a_interp.y)


def test_config3_spa_saved_trajectories_full_size(torch_cuda):
    from trajectories import _engine as eng
    from trajectories.distributions import CeNTREXVelocityDistribution, GaussianPositionDistribution
    from trajectories.trajectory_simulator import TrajectorySimulator

    bl = spa_beamline()
    seed = 13
    sim = TrajectorySimulator(seed=seed)
    t0 = time.perf_counter()
    sim.run_simulation(bl, "SPA", N_traj=int(1e9), apertures_of_interest=["Detected"], n_jobs=9,
                       xdist=GaussianPositionDistribution())
    seconds = time.perf_counter() - t0
    n_run = 900 * int(1e9 / 900)
    c = sim.counter.counter_dict
    assert sum(c.values()) == n_run
    mols = sim.result.molecules
    print(f"configs[3]: {n_run} molecules, {len(mols)} detected trajectories saved in {seconds:.2f} s = {n_run / seconds:.3g} molecules/s")
    assert len(mols) == c["Detected"] and abs(c["Detected"] / n_run - 3.1e-4) < 2e-5        # SURVEY.md 3.4 [probe]
    assert sim.result.molecule_offset == 0 and sim.result.n_molecules_total == len(mols)
    for m in mols[:: max(1, len(mols) // 500)]:
        assert m.trajectory.x.shape == (19, 3) and m.alive and m.aperture_hit == "Detected"   # 1 + 2 x 9 rows

    # a 1e7 sub-range of the same run against the oracle: same seed, same global indices
    first, n = 420_000_000, 10_000_000
    src_dev = eng.make_source(CeNTREXVelocityDistribution(), GaussianPositionDistribution())
    prop = eng.Propagator(bl.elements, 0)
    prop.reset()
    mask = prop.flat.save_mask(["Detected"])
    r = prop.propagate_philox(src_dev, seed, first, n, save_mask=mask)
    torch_cuda.cuda.synchronize()
    src = oracle.make_source(CeNTREXVelocityDistribution(), GaussianPositionDistribution())
    want = oracle.run(bl.elements, src, seed, first, n, n_threads=oracle.host_cores())
    got = {nm: int(v) for nm, v in zip(prop.flat.fate_names, r.counters.cpu().tolist()) if v > 0}
    counter_close(got, want["fate_names"], want["counters"], n)
    # the run's saved molecules inside that range are exactly the ones this launch selected, in global order
    sel = r.saved_index.cpu().numpy()
    assert np.all(np.diff(sel) > 0) and sel.min() >= first and sel.max() < first + n
    ic = prop.draw(src_dev, seed, index=r.saved_index).cpu().numpy()
    x0 = np.array([m.trajectory.x[0] for m in mols])
    # locate the block by its first molecule (trajectories are in global-index order)
    k0 = int(np.flatnonzero((x0[:, 0] == ic[0, 0]) & (x0[:, 1] == ic[1, 0]))[0])
    np.testing.assert_array_equal(x0[k0:k0 + sel.size].T, ic[0:3])
    # and their rows are what the oracle computes from those initial conditions
    rows = oracle.propagate(bl.elements, ic[:, :300], want_rows=True)["rows"]
    for j in range(min(300, sel.size)):
        m = mols[k0 + j]
        got_rows = np.concatenate([m.trajectory.x, m.trajectory.v, m.trajectory.a, m.trajectory.t[:, None]], axis=1)
        assert np.max(np.abs(got_rows - rows[j, :19]) / np.maximum(np.abs(rows[j, :19]), 1e-12)) < 1e-12
