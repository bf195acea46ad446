"""Right-sized lens queues (cmt_outputs_t.queue_capacity): a launch whose queue is smaller than the launch gives the
results of a full-size queue as long as it does not overflow, reports every molecule it had to drop otherwise, and
run_simulation sizes its queues from a pilot launch and falls back to full-size queues on overflow."""
import numpy as np
import pytest

from tests.beamlines import lens_beamline, lens_table, standard_ics

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


def test_small_queue_same_result_and_overflow_is_counted(torch_cuda):
    from trajectories import _engine as eng

    bl = lens_beamline(lens_table())
    n = 400_000
    ic = torch_cuda.from_numpy(standard_ics(n, 4)).cuda()
    full = eng.Propagator(bl.elements, 0)
    a = full.propagate_ic(ic, want_fate=True, want_final=True)
    entries = int(a.work[3].item())
    assert 1000 < entries < n // 64 and int(a.work[6].item()) == 0

    sized = eng.Propagator(bl.elements, 0)
    sized.entry_fraction = entries / n
    cap = sized.queue_capacity(n)
    assert entries < cap <= max(n // 64, 2 * entries + 4096)
    assert sized.dev.workspace_bytes(cap) < full.dev.workspace_bytes(n) // 32
    b = sized.propagate_ic(ic, want_fate=True, want_final=True)
    assert sized._ws[sized.n_slots].numel() == sized.dev.workspace_bytes(cap)
    np.testing.assert_array_equal(a.fate.cpu().numpy(), b.fate.cpu().numpy())
    np.testing.assert_array_equal(a.final.cpu().numpy(), b.final.cpu().numpy())
    np.testing.assert_array_equal(a.counters.cpu().numpy(), b.counters.cpu().numpy())
    np.testing.assert_array_equal(a.work.cpu().numpy(), b.work.cpu().numpy())

    # a queue that is too small: every molecule is either counted under a fate or reported as dropped
    tiny = eng.Propagator(bl.elements, 0)
    tiny.entry_fraction = 0.0
    tiny.queue_capacity = lambda m: 500
    c = tiny.propagate_ic(ic, want_fate=False)
    dropped = tiny.queue_overflow()
    assert dropped == entries - 500
    assert int(c.counters.sum().item()) + dropped == n
    front = bl.elements[:3]
    names = tiny.flat.fate_names
    for e in front:                                   # what happened before the lens is untouched
        k = names.index(e.name)
        assert int(c.counters[k].item()) == int(a.counters[k].item())


def test_queue_capacity_needs_work_counters(torch_cuda, cuda_lib):
    import ctypes as C

    from trajectories import _engine as eng
    from trajectories import _native as nat

    prop = eng.Propagator(lens_beamline(lens_table()).elements, 0)
    ic = torch_cuda.from_numpy(standard_ics(1000, 1)).cuda()
    O = nat.Outputs()
    O.counters = prop.counters.data_ptr()
    O.queue_capacity = 100
    ws = torch_cuda.empty(prop.dev.workspace_bytes(1000), dtype=torch_cuda.uint8, device="cuda")
    rc = cuda_lib.cmt_propagate_ic(prop.dev.handle, 1000, 0, ic.data_ptr(), ic.stride(0), C.byref(O), ws.data_ptr(),
                                   ws.numel(), None)
    assert rc != 0 and b"queue_capacity" in cuda_lib.cmt_last_error()
    O.queue_capacity = -1
    assert cuda_lib.cmt_propagate_ic(prop.dev.handle, 1000, 0, ic.data_ptr(), ic.stride(0), C.byref(O), ws.data_ptr(),
                                     ws.numel(), None) != 0


def test_run_simulation_pilot_and_fallback(torch_cuda, monkeypatch):
    from trajectories import _engine as eng
    from trajectories.trajectory_simulator import TrajectorySimulator

    bl = lens_beamline(lens_table())
    n, seed = 30_000_000, 17
    assert n > 8 * eng.PILOT_MOLECULES
    torch_cuda.cuda.reset_peak_memory_stats()
    base = torch_cuda.cuda.memory_allocated()
    sim = TrajectorySimulator(seed=seed)
    sim.run_simulation(bl, "pilot", N_traj=n, n_jobs=10)
    peak_sized = torch_cuda.cuda.max_memory_allocated() - base
    assert sum(sim.counter.counter_dict.values()) == n and sim.last_work[6] == 0

    # the same run with queues that hold every molecule (no pilot)
    with monkeypatch.context() as mp:
        mp.setattr(eng, "PILOT_MOLECULES", 1 << 40)
        torch_cuda.cuda.reset_peak_memory_stats()
        ref = TrajectorySimulator(seed=seed)
        ref.run_simulation(bl, "full", N_traj=n, n_jobs=10)
        peak_full = torch_cuda.cuda.max_memory_allocated() - base
    assert ref.counter.counter_dict == sim.counter.counter_dict
    np.testing.assert_array_equal(ref.last_work[:5], sim.last_work[:5])
    assert peak_sized < peak_full / 4, (peak_sized, peak_full)

    # queues sized far too small after the pilot: the run is repeated with full-size queues, same Counter
    with monkeypatch.context() as mp:
        mp.setattr(eng.Propagator, "queue_capacity", lambda self, m: int(m) if self.entry_fraction is None else 1000)
        again = TrajectorySimulator(seed=seed)
        again.run_simulation(bl, "fallback", N_traj=n, n_jobs=10)
    assert again.counter.counter_dict == sim.counter.counter_dict and again.last_work[6] == 0

    # with saved trajectories: the pilot's molecules come first, the list is the one of the unsized run
    small = TrajectorySimulator(seed=seed)
    small.run_simulation(bl, "saved", N_traj=20_000_000, apertures_of_interest=["Detected"], n_jobs=10)
    with monkeypatch.context() as mp:
        mp.setattr(eng, "PILOT_MOLECULES", 1 << 40)
        want = TrajectorySimulator(seed=seed)
        want.run_simulation(bl, "saved full", N_traj=20_000_000, apertures_of_interest=["Detected"], n_jobs=10)
    assert small.counter.counter_dict == want.counter.counter_dict
    assert len(small.result.molecules) == len(want.result.molecules) == want.counter.counter_dict["Detected"]
    for ma, mb in zip(small.result.molecules[::97], want.result.molecules[::97]):
        np.testing.assert_array_equal(ma.trajectory.x, mb.trajectory.x)


def test_run_sweep_pilot_and_fallback(torch_cuda, monkeypatch):
    """run_sweep sizes the lens queues of every point from one small pilot launch of the first point; the points are
    what they are with full-size queues, and queues that overflow all the same send the sweep through full-size ones."""
    from trajectories import _engine as eng
    from trajectories.trajectory_simulator import TrajectorySimulator

    bl = lens_beamline()
    states, volts, n, seed = [(2, 0), (1, 1)], [24e3, 30e3], 4_000_000, 5
    assert n > 4 * eng.SWEEP_PILOT_MOLECULES
    sized = TrajectorySimulator(seed=seed).run_sweep(bl, states, volts, N_traj=n, n_jobs=10)
    with monkeypatch.context() as mp:
        mp.setattr(eng, "SWEEP_PILOT_MOLECULES", 1 << 40)           # no pilot: every queue holds every molecule
        full = TrajectorySimulator(seed=seed).run_sweep(bl, states, volts, N_traj=n, n_jobs=10)
    with monkeypatch.context() as mp:
        mp.setattr(eng.Propagator, "queue_capacity", lambda self, m: int(m) if self.entry_fraction is None else 500)
        sim = TrajectorySimulator(seed=seed)
        again = sim.run_sweep(bl, states, volts, N_traj=n, n_jobs=10)
        assert sim.last_work[6] == 0
    assert set(sized) == set(full) == set(again) and len(sized) == 4
    for key in full:
        want = full[key].counter.counter_dict
        assert sum(want.values()) == n
        assert sized[key].counter.counter_dict == want and again[key].counter.counter_dict == want
