"""Plane-crossing probe (cmt_plane_crossings): the device form of the reference's
post_processing.find_radial_pos_dist / find_vel_dist (post_processing.py:20-140).

Checked against (1) what the unmodified reference returned (tests/golden/plane_crossings.npz),
(2) the host post-processing applied to the oracle's rows on larger seeded sets, and (3) the
public API: run_simulation + find_*_dist on saved trajectories vs plane_distributions.

Which molecules count at a plane is exact.  Values: bit-identical except where NumPy's `dt ** 2`
(libm pow) and the exact square differ in the last bit; asserted <= 1e-12 relative (+1e-15 absolute).
"""
import json
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle
from tests.beamlines import lens_beamline, lens_table, spa_beamline, standard_ics


@pytest.fixture(scope="module")
def torch_cuda(cuda_lib):
    import torch

    assert torch.cuda.is_available(), "GPU tests need a B200"
    return torch


def close(a, b, rtol=1e-12, atol=1e-15):
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= rtol * np.abs(b) + atol))


def same_bits(a, b):
    return float((np.ascontiguousarray(a).view(np.int64) == np.ascontiguousarray(b).view(np.int64)).mean()) if a.size else 1.0


def device_crossings(torch, beamline, ic, planes, math="exact", select=None, select_base=0):
    from trajectories import _engine as eng

    prop = eng.Propagator(beamline.elements, 0, math=math)
    out, valid, fate = prop.plane_crossings(torch.from_numpy(np.ascontiguousarray(ic)).cuda(), planes,
                                            select=select, select_base=select_base)
    torch.cuda.synchronize()
    return out.cpu().numpy(), valid.cpu().numpy(), fate.cpu().numpy()


def host_crossings(beamline, ic, planes):
    """find_radial_pos_dist / find_vel_dist of the package (bit-exact with the reference, see
    test_oracle_golden.py::test_plane_crossings_host) on the oracle's rows; per molecule, NaN where not reached."""
    from trajectories.molecule import Molecule
    from trajectories.post_processing import state_at_plane

    res = oracle.propagate(beamline.elements, ic, want_rows=True)
    n = ic.shape[1]
    out = np.full((len(planes), 5, n), np.nan)
    for i in range(n):
        mol = Molecule.from_rows(res["rows"][i, : res["n_rows"][i]], "", True)
        for p, z in enumerate(planes):
            st = state_at_plane(mol, float(z))
            if st is not None:
                out[p, 0:2, i], out[p, 2:5, i] = st[0][:2], st[1]
    return out, res["fate"]


def test_golden_planes(torch_cuda, golden_dir):
    g = np.load(golden_dir / "plane_crossings.npz")
    bl = lens_beamline((g["table_r"], g["table_a"]))
    names = list(g["fate_names"])
    planes = g["planes"]
    out, valid, fate = device_crossings(torch_cuda, bl, g["ic"], planes)
    np.testing.assert_array_equal(fate, g["fate"])
    filters = json.loads(str(g["filters"]))
    bits = []
    for p in range(len(planes)):
        for f, elements in enumerate(filters):
            keep = valid[p].copy()
            if elements is not None:
                keep &= np.isin(fate, [names.index(e) for e in elements])
            want_xy, want_v = g[f"xy_{p}_{f}"], g[f"v_{p}_{f}"]
            assert keep.sum() == want_xy.shape[0], (planes[p], elements)       # the same molecules count
            if want_xy.shape[0] == 0:
                continue
            got_xy, got_v = out[p, 0:2][:, keep].T, out[p, 2:5][:, keep].T
            assert close(got_xy, want_xy) and close(got_v, want_v), (planes[p], elements)
            bits.append(same_bits(np.concatenate([got_xy, got_v], axis=1), np.concatenate([want_xy, want_v], axis=1)))
    assert min(bits) > 0.98 and np.mean(bits) > 0.995


@pytest.mark.parametrize("n,seed,sigma", [(3000, 21, 3.0), (20000, 22, 39.5), (65, 23, 2.0)])
def test_planes_against_oracle_rows(torch_cuda, n, seed, sigma):
    bl = lens_beamline(lens_table())
    ic = standard_ics(n, seed, sigma)
    lens = bl.elements[3]
    # unsorted on purpose, with duplicates, element planes, and more than CMT_MAX_PLANES entries
    planes = [2.0, 0.001, lens.z0, 1.3, 1.3, lens.z1, 0.00635, 0.1, 0.2, 0.6, 1.1, 1.45, 1.7, 2.43, 3.0, 4.4, 5.43,
              bl.elements[5].z0, bl.elements[5].z1, 7.0, bl.elements[1].z1]
    assert len(planes) > 16
    want, want_fate = host_crossings(bl, ic, planes)
    out, valid, fate = device_crossings(torch_cuda, bl, ic, planes)
    np.testing.assert_array_equal(fate, want_fate)
    np.testing.assert_array_equal(valid, ~np.isnan(want[:, 0]))
    for p in range(len(planes)):
        k = valid[p]
        assert close(out[p][:, k], want[p][:, k]), planes[p]
    assert same_bits(out[valid[:, None, :].repeat(5, 1)], want[valid[:, None, :].repeat(5, 1)]) > 0.995
    np.testing.assert_array_equal(out[4], out[3])                      # duplicate plane, same answer
    assert not valid[planes.index(7.0)].any()                          # beyond the last element: nobody gets there


def test_planes_spa_and_select(torch_cuda):
    bl = spa_beamline()
    rng = np.random.default_rng(31)
    n = 5000
    ic = np.empty((6, n))
    ic[0], ic[1], ic[2] = rng.normal(0, 1e-3, n), rng.normal(0, 1e-3, n), 0.00635
    ic[3], ic[4], ic[5] = rng.normal(0, 1.5, n), rng.normal(0, 1.5, n), rng.normal(184, 16, n)
    planes = [0.3, 0.9, bl.elements[-1].z0, 1.25]
    want, _ = host_crossings(bl, ic, planes)
    out, valid, _ = device_crossings(torch_cuda, bl, ic, planes)
    np.testing.assert_array_equal(valid, ~np.isnan(want[:, 0]))
    assert valid[2].sum() > 50
    for p in range(len(planes)):
        assert close(out[p][:, valid[p]], want[p][:, valid[p]])
    # select: gather columns by global index
    idx = np.sort(rng.choice(n, 700, replace=False)).astype(np.int64)
    sel = torch_cuda.from_numpy(idx + 10**9).cuda()
    out2, valid2, _ = device_crossings(torch_cuda, bl, ic, planes, select=sel, select_base=10**9)
    np.testing.assert_array_equal(valid2, valid[:, idx])
    np.testing.assert_array_equal(out2[valid2[:, None, :].repeat(5, 1)], out[:, :, idx][valid2[:, None, :].repeat(5, 1)])


def test_plane_distributions_api(torch_cuda):
    """run_simulation + find_*_dist on saved trajectories == plane_distributions, same seed."""
    from trajectories.post_processing import find_radial_pos_dist, find_vel_dist
    from trajectories.trajectory_simulator import TrajectorySimulator

    bl = lens_beamline(lens_table())
    keep = ["Detected", "Field plates"]
    sim = TrajectorySimulator(seed=77)
    sim.run_simulation(bl, "a", N_traj=400000, apertures_of_interest=keep, n_jobs=4)
    assert len(sim.result.molecules) > 50
    zs = [1.3, 2.0, 5.5]
    sim2 = TrajectorySimulator(seed=77)
    pairs = sim2.plane_distributions(bl, zs, elements=keep, N_traj=400000, n_jobs=4)
    assert sim2.counter.counter_dict == sim.counter.counter_dict
    for z, (xy, v) in zip(zs, pairs):
        want_xy, want_v = find_radial_pos_dist(sim.result, z, keep), find_vel_dist(sim.result, z, keep)
        assert xy.shape == want_xy.shape and xy.shape[0] > 10
        assert close(xy, want_xy) and close(v, want_v)
    xy1, v1 = sim2.plane_distributions(bl, 2.0, elements=keep, N_traj=400000, n_jobs=4)      # scalar z
    np.testing.assert_array_equal(xy1, pairs[1][0])
    # no element filter: everybody who reaches the plane; Counter from the same pass
    sim3 = TrajectorySimulator(seed=77)
    xy_all, v_all = sim3.plane_distributions(bl, 0.5, N_traj=400000, n_jobs=4)
    assert sim3.counter.counter_dict == sim.counter.counter_dict
    c = sim.counter.counter_dict
    reached = sum(c.values()) - c.get("4K shield", 0) - c.get("40K shield", 0) - c.get("BB exit", 0)
    assert xy_all.shape == (reached, 2) and v_all.shape == (reached, 3)
    assert np.all(np.hypot(xy_all[:, 0], xy_all[:, 1]) < 0.2)
    # a custom Distribution goes through the host-draw path
    from trajectories.distributions import CeNTREXVelocityDistribution, Distribution

    class Narrow(Distribution):
        def draw(self, n):
            return CeNTREXVelocityDistribution(sigmax=2, sigmay=2).draw(n)

        def save_to_hdf(self, *a, **k):
            pass

    np.random.seed(5)
    sim4 = TrajectorySimulator()
    sim4.run_simulation(bl, "b", vdist=Narrow(), N_traj=20000, apertures_of_interest=["Detected"], n_jobs=1)
    np.random.seed(5)
    sim5 = TrajectorySimulator()
    xy5, v5 = sim5.plane_distributions(bl, 1.4, elements=["Detected"], vdist=Narrow(), N_traj=20000, n_jobs=1)
    assert sim5.counter.counter_dict == sim4.counter.counter_dict
    assert close(xy5, find_radial_pos_dist(sim4.result, 1.4, ["Detected"])) and xy5.shape[0] > 20
    assert close(v5, find_vel_dist(sim4.result, 1.4, ["Detected"]))


def test_planes_contracted_mode(torch_cuda):
    bl = lens_beamline(lens_table())
    ic = standard_ics(20000, 41, 3.0)
    planes = [0.5, 1.3, 2.0, 5.0]
    a, va, fa = device_crossings(torch_cuda, bl, ic, planes)
    b, vb, fb = device_crossings(torch_cuda, bl, ic, planes, math="contracted")
    assert (fa != fb).mean() < 1e-4 and (va != vb).mean() < 1e-4
    both = (va & vb)[:, None, :].repeat(5, 1)
    assert np.all(np.abs(a[both] - b[both]) <= 1e-9 * np.abs(a[both]) + 1e-12)


def test_plane_argument_checks(torch_cuda, cuda_lib):
    import ctypes as C

    from trajectories import _engine as eng
    from trajectories._native import NativeError

    torch = torch_cuda
    bl = lens_beamline(lens_table())
    prop = eng.Propagator(bl.elements, 0)
    ic = torch.from_numpy(standard_ics(10, 1)).cuda()
    out = torch.zeros((2, 5, 10), dtype=torch.float64, device="cuda")
    valid = torch.zeros((2, 10), dtype=torch.uint8, device="cuda")
    z = (C.c_double * 2)(2.0, 1.0)                                    # descending
    rc = cuda_lib.cmt_plane_crossings(prop.dev.handle, 10, ic.data_ptr(), 6, ic.stride(0), None, 0, z, 2,
                                      out.data_ptr(), 10, valid.data_ptr(), None, None)
    assert rc < 0 and b"ascending" in cuda_lib.cmt_last_error()
    rc = cuda_lib.cmt_plane_crossings(prop.dev.handle, 10, ic.data_ptr(), 6, ic.stride(0), None, 0, z, 17,
                                      out.data_ptr(), 10, valid.data_ptr(), None, None)
    assert rc < 0
    with pytest.raises(ValueError):
        prop.plane_crossings(ic, [])
    with pytest.raises(ValueError):
        prop.plane_crossings(ic, [float("nan")])
    o, v, f = prop.plane_crossings(ic[:, :0], [1.0])                   # empty input
    assert o.shape == (1, 5, 0) and v.shape == (1, 0) and f.shape == (0,)
