"""User-defined BeamlineElement subclasses in a GPU run (the reference's plugin API, apertures.py:22-54).

The element under test is the reference's own CircularAperture.propagate_through (apertures.py:92-115) written
against the Molecule API, as a user would write an element of their own.  Put in place of a built-in aperture it
must change nothing: same Counter, same saved trajectories as the all-CUDA run on the same sample."""
from dataclasses import dataclass

import numpy as np
import pytest

from tests.beamlines import lens_beamline, lens_table, spa_beamline, standard_ics

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


def user_elements():
    from trajectories.beamline_elements.apertures import BeamlineElement

    @dataclass
    class PyCircular(BeamlineElement):
        """apertures.py:92-115, line for line against the Molecule API."""
        d: float = 0.0254
        calls: int = 0

        def propagate_through(self, molecule):
            self.calls += 1
            for z in [self.z0, self.z1]:
                delta_t = (z - molecule.x()[2]) / molecule.v()[2]
                molecule.update_trajectory(delta_t)
                rho = np.sqrt(np.sum(molecule.x()[:2] ** 2))
                if rho > self.d / 2:
                    molecule.set_dead()
                    molecule.set_aperture_hit(self.name)
                    return

        def N_steps(self):
            return 2

    @dataclass
    class BatchCircular(BeamlineElement):
        """The same test, vectorised over the molecules that reach the element (this build's batch protocol)."""
        d: float = 0.0254

        def propagate_through(self, molecule):  # pragma: no cover - the batch form is preferred
            raise AssertionError("the batch form must be used")

        def propagate_through_batch(self, rows):
            k = rows.shape[0]
            alive = np.ones(k, dtype=bool)
            out = rows.copy()
            for z in [self.z0, self.z1]:
                x, v, a, t = out[:, 0:3], out[:, 3:6], out[:, 6:9], out[:, 9]
                dt = ((z - x[:, 2]) / v[:, 2])[:, None]
                nx = x + v * dt + a * dt ** 2 / 2
                nv = v + a * dt
                new = np.concatenate([nx, nv, np.broadcast_to([0.0, -9.80665, 0.0], (k, 3)), t[:, None] + dt], axis=1)
                out = np.where(alive[:, None], new, out)
                rho = np.sqrt(np.sum(out[:, :2] ** 2, axis=1))
                alive &= ~(rho > self.d / 2)
            return alive, out, self.name

        def N_steps(self):
            return 2

    @dataclass
    class Scatterer(BeamlineElement):
        """Kills a random third of what reaches it and kicks the rest sideways: stochastic, with a fate of its own."""
        seed: int = 5

        def __post_init__(self):
            super().__post_init__()
            self.rng = np.random.default_rng(self.seed)

        def propagate_through(self, molecule):
            molecule.update_trajectory((self.z0 - molecule.x()[2]) / molecule.v()[2])
            if self.rng.random() < 1 / 3:
                molecule.set_dead()
                molecule.set_aperture_hit("scattered away")
                return
            kick = np.array((self.rng.normal(0, 0.05), 0.0, 0.0))
            tr = molecule.trajectory
            tr.update(molecule.x(), molecule.v() + kick, molecule.a(), molecule.t())

        def N_steps(self):
            return 2

    return PyCircular, BatchCircular, Scatterer


def replaced(beamline, name, cls, **kw):
    """The beamline with the built-in circular aperture `name` replaced by a user element of the same geometry."""
    from trajectories.beamline import Beamline

    out = []
    for e in beamline.elements:
        out.append(cls(name=e.name, z0=e.z0, L=e.L, d=e.d, **kw) if e.name == name else e)
    return Beamline(out)


@pytest.mark.parametrize("which", ["4K shield", "BB exit"])
def test_user_element_in_place_of_a_builtin_changes_nothing(torch_cuda, which):
    from trajectories.trajectory_simulator import TrajectorySimulator

    PyCircular, BatchCircular, _ = user_elements()
    n, seed = 200_000, 11
    ref = TrajectorySimulator(seed=seed)
    ref.run_simulation(lens_beamline(lens_table()), "cuda", N_traj=n, apertures_of_interest=["Detected", "Field plates"], n_jobs=10)
    assert ref.counter.counter_dict["Detected"] > 20

    for cls in (PyCircular, BatchCircular):
        bl = replaced(lens_beamline(lens_table()), which, cls)
        sim = TrajectorySimulator(seed=seed)
        sim.run_simulation(bl, "hybrid", N_traj=n, apertures_of_interest=["Detected", "Field plates"], n_jobs=10)
        assert sim.counter.counter_dict == ref.counter.counter_dict, cls.__name__
        a, b = ref.result.molecules, sim.result.molecules
        assert len(a) == len(b) == ref.counter.counter_dict["Detected"] + ref.counter.counter_dict["Field plates"]
        for ma, mb in zip(a, b):
            assert ma.aperture_hit == mb.aperture_hit and ma.alive == mb.alive
            if cls is BatchCircular:
                # the batch form records one row for the element instead of two: compare the common end
                np.testing.assert_allclose(ma.trajectory.x[-1], mb.trajectory.x[-1], rtol=1e-12, atol=1e-18)
                continue
            assert ma.trajectory.x.shape == mb.trajectory.x.shape
            for key in ("x", "v", "a", "t"):
                np.testing.assert_allclose(getattr(ma.trajectory, key), getattr(mb.trajectory, key), rtol=1e-12, atol=1e-18)
        if cls is PyCircular:
            el = bl.find_element(which)
            # called once per molecule that reached it -- never a second time for the saved ones
            ahead = [e.name for e in bl.elements[:bl.elements.index(el)]]
            assert el.calls == n - sum(ref.counter.counter_dict.get(k, 0) for k in ahead)


def test_user_element_first_last_and_alone(torch_cuda):
    from trajectories.beamline import Beamline
    from trajectories.trajectory_simulator import TrajectorySimulator

    PyCircular, _, _ = user_elements()
    n, seed = 50_000, 3
    want = TrajectorySimulator(seed=seed)
    want.run_simulation(spa_beamline(), "cuda", N_traj=n, apertures_of_interest=["Detected"], n_jobs=5)
    for name in ("4K shield", "DR entrance"):                       # first element; a late one
        sim = TrajectorySimulator(seed=seed)
        sim.run_simulation(replaced(spa_beamline(), name, PyCircular), "hybrid", N_traj=n,
                           apertures_of_interest=["Detected"], n_jobs=5)
        assert sim.counter.counter_dict == want.counter.counter_dict
        assert len(sim.result.molecules) == len(want.result.molecules)
        for ma, mb in zip(want.result.molecules, sim.result.molecules):
            np.testing.assert_allclose(ma.trajectory.x, mb.trajectory.x, rtol=1e-12, atol=1e-18)
            np.testing.assert_allclose(ma.trajectory.t, mb.trajectory.t, rtol=1e-12, atol=1e-18)
    # a beamline of user elements only
    only = Beamline([PyCircular(name=e.name, z0=e.z0, L=e.L, d=e.d) for e in spa_beamline().elements[:3]])
    sim = TrajectorySimulator(seed=seed)
    sim.run_simulation(only, "host only", N_traj=5_000, n_jobs=5)
    front = Beamline(spa_beamline().elements[:3])
    ref = TrajectorySimulator(seed=seed)
    ref.run_simulation(front, "cuda", N_traj=5_000, n_jobs=5)
    assert sim.counter.counter_dict == ref.counter.counter_dict


def test_stochastic_element_is_called_once_per_molecule(torch_cuda):
    from trajectories.beamline import Beamline
    from trajectories.trajectory_simulator import TrajectorySimulator

    _, _, Scatterer = user_elements()
    bl = lens_beamline(lens_table())
    els = list(bl.elements)
    lens = bl.find_element("ES lens")
    els.append(Scatterer(name="gas cell", z0=lens.z1 + 0.2, L=0.01))
    hybrid = Beamline(els)
    sim = TrajectorySimulator(seed=21)
    sim.run_simulation(hybrid, "scatter", N_traj=400_000, apertures_of_interest=["Detected", "scattered away"], n_jobs=10)
    c = sim.counter.counter_dict
    assert sum(c.values()) == 400_000
    through_lens = 400_000 - sum(c.get(k, 0) for k in ("4K shield", "40K shield", "BB exit", "Lens entrance", "Inside lens"))
    assert through_lens > 100
    assert abs(c["scattered away"] - through_lens / 3) < 5 * np.sqrt(through_lens)       # a third, binomially
    saved = sim.result.molecules
    assert len(saved) == c["Detected"] + c["scattered away"]
    gas = hybrid.find_element("gas cell")
    for m in saved:
        z = m.trajectory.x[:, 2]
        assert np.all(np.diff(m.trajectory.t) >= 0) and np.isfinite(m.trajectory.x).all()
        if m.aperture_hit == "scattered away":
            assert not m.alive and abs(z[-1] - gas.z0) < 1e-12
        else:
            # the kick row repeats the position at the gas cell; rows of the device segment behind it follow
            k = int(np.argmin(np.abs(z - gas.z0)))
            assert z[k + 1] == z[k] and m.trajectory.v[k + 1, 0] != m.trajectory.v[k, 0]
            assert z[-1] > gas.z1
    # the counts of the elements before the scatterer are those of the plain CUDA run on the same seed
    plain = TrajectorySimulator(seed=21)
    plain.run_simulation(lens_beamline(lens_table()), "plain", N_traj=400_000, n_jobs=10)
    for k in ("4K shield", "40K shield", "BB exit", "Lens entrance", "Inside lens"):
        assert c.get(k, 0) == plain.counter.counter_dict.get(k, 0)


def test_single_molecule_api_on_a_mixed_beamline(torch_cuda):
    from trajectories.molecule import Molecule

    PyCircular, _, _ = user_elements()
    ic = standard_ics(64, 9, sigma_perp=4.0)
    plain, mixed = lens_beamline(lens_table()), replaced(lens_beamline(lens_table()), "BB exit", PyCircular)
    for j in range(0, 64, 7):
        a, b = Molecule(), Molecule()
        a.init_trajectory(plain, ic[0:3, j], ic[3:6, j])
        b.init_trajectory(mixed, ic[0:3, j], ic[3:6, j])
        plain.propagate_through(a)
        mixed.propagate_through(b)
        assert a.aperture_hit == b.aperture_hit and a.alive == b.alive
        np.testing.assert_allclose(a.trajectory.x, b.trajectory.x, rtol=1e-12, atol=1e-18)


def test_resume_entry_point(torch_cuda, cuda_lib):
    """cmt_resume against cmt_trajectories on the same 10-component states: same fate, last row = last trajectory row."""
    from trajectories import _engine as eng
    from trajectories import _hybrid

    bl = lens_beamline(lens_table())
    prop = eng.Propagator(bl.elements[3:], 0)                     # lens, field plates, DR aperture
    ic = standard_ics(4000, 5, sigma_perp=3.0)
    state = np.zeros((10, 4000))
    state[0:6] = ic
    state[2] = 0.9                                               # already past the front apertures
    state[6], state[7], state[9] = 0.0, -9.80665, 4.2e-3
    dev = torch_cuda.from_numpy(state).cuda()
    fate, last = _hybrid._resume(prop, dev)
    rows, off, fate2 = prop.trajectories(dev)
    np.testing.assert_array_equal(fate.cpu().numpy(), fate2)
    got = last.cpu().numpy().T
    want = np.stack([rows[off[k + 1] - 1] for k in range(4000)])
    np.testing.assert_array_equal(got, want)
    assert len(set(fate2.tolist())) >= 3
    assert cuda_lib.cmt_resume(None, 1, None, 1, None, 1, None, None, None) == -1
