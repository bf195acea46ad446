"""World-size-2 test of the multi-GPU host logic on CPU (gloo): shards partition
the global index range and the Counter all-reduce sums the per-rank histograms.
The per-rank histograms are produced by the oracle here (no GPU); on the GPU box
the same code path runs with NCCL tensors."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    from oracle import oracle
    from tests.beamlines import lens_beamline, lens_table
    from trajectories import _engine as eng
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution

    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert eng.dist_info() == (rank, world)
    lo, hi = eng.shard_range(total, rank, world)
    bl = lens_beamline(lens_table())
    src = oracle.make_source(CeNTREXVelocityDistribution(), CeNTREXPositionDistribution())
    local = oracle.run(bl.elements, src, seed=4, first=lo, n=hi - lo)["counters"]
    t = torch.from_numpy(local.copy())
    eng.allreduce_counts(t)
    np.save(Path(out_dir) / f"rank{rank}.npy", np.stack([local, t.numpy()]))
    dist.destroy_process_group()


def test_shard_and_allreduce_world2(tmp_path):
    import torch.multiprocessing as mp

    from oracle import oracle
    from tests.beamlines import lens_beamline, lens_table
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    total = 200_001
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    np.testing.assert_array_equal(r0[1], r1[1])                 # both ranks hold the merged Counter
    np.testing.assert_array_equal(r0[0] + r1[0], r0[1])
    bl = lens_beamline(lens_table())
    src = oracle.make_source(CeNTREXVelocityDistribution(), CeNTREXPositionDistribution())
    whole = oracle.run(bl.elements, src, seed=4, first=0, n=total)["counters"]
    np.testing.assert_array_equal(r0[1], whole)                 # GPU-count invariance of the result
    assert whole.sum() == total


# ---------------------------------------------------------------------------
# saved trajectories across ranks: the gather of the saved counts, global molecule numbering, one merged file
# ---------------------------------------------------------------------------
def _fake_molecules(lo, hi):
    """Molecule k of a pretend run: k % 4 + 2 rows whose values encode k."""
    from trajectories.molecule import Molecule

    out = []
    for k in range(lo, hi):
        rows = (np.arange((k % 4 + 2) * 10, dtype=float).reshape(-1, 10) + 1000.0 * k)
        out.append(Molecule.from_rows(rows, "Detected" if k % 3 else "Inside lens", bool(k % 3)))
    return out


def _result(mols, offset=0, total=None):
    from tests.beamlines import apertures_beamline
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution
    from trajectories.trajectory_simulator import Counter, SimulationResult

    c = Counter()
    c.increment_counter("Detected", 7)
    return SimulationResult(c, apertures_beamline(), CeNTREXPositionDistribution(), CeNTREXVelocityDistribution(), mols,
                            offset, total)


def _save_worker(rank, world, port, out_dir, packed):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.modules["h5py"] = None                      # the built-in writer, as on the GPU box
    import torch.distributed as dist

    from trajectories import _engine as eng

    dist.init_process_group("gloo", rank=rank, world_size=world)
    cuts = [0, 5, 5, 14][: world + 1] if world == 3 else [0, 5, 14]       # rank 1 of 3 saved nothing
    mols = _fake_molecules(cuts[rank], cuts[rank + 1])
    counts = eng.gather_counts(len(mols))
    assert counts == [cuts[r + 1] - cuts[r] for r in range(world)]
    res = _result(mols, sum(counts[:rank]), sum(counts))
    res.save_to_hdf(Path(out_dir) / "merged.hdf", "run/a", packed=packed)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,packed", [(2, False), (3, False), (2, True)])
def test_ranks_write_one_file_numbered_like_a_single_run(tmp_path, monkeypatch, world, packed):
    """SimulationResult.save_to_hdf under torch.distributed: metadata once, molecules of rank r numbered from the
    count saved by ranks < r (reference numbering trajectory_simulator.py:251-254 over the flat list of :86-91)."""
    import torch.multiprocessing as mp

    from trajectories import _minih5, utils

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_save_worker, args=(world, port, str(tmp_path), packed), nprocs=world, join=True)
    monkeypatch.setitem(sys.modules, "h5py", None)
    single = tmp_path / "single.hdf"
    _result(_fake_molecules(0, 14)).save_to_hdf(single, "run/a", packed=packed)
    with _minih5.File(tmp_path / "merged.hdf", "r") as m, _minih5.File(single, "r") as s1:
        assert m["run/a"].keys() == s1["run/a"].keys()
        assert dict(m["run/a/counter"].attrs.items()) == dict(s1["run/a/counter"].attrs.items())
        assert m["run/a/beamline"].keys() == s1["run/a/beamline"].keys()
        if not packed:
            assert m["run/a/trajectories"].keys() == s1["run/a/trajectories"].keys()
            assert len(m["run/a/trajectories"]) == 14
            for name in s1["run/a/trajectories"].keys():
                a, b = m["run/a/trajectories"][name], s1["run/a/trajectories"][name]
                for key in ("x", "v", "a", "t"):
                    np.testing.assert_array_equal(a[key][()], b[key][()])
                assert dict(a.attrs.items()) == dict(b.attrs.items())
    merged = utils.import_trajectories_from_hdf(tmp_path / "merged.hdf", "run/a")
    alone = utils.import_trajectories_from_hdf(single, "run/a")
    assert len(merged) == len(alone) == 14
    for a, b in zip(merged, alone):
        np.testing.assert_array_equal(a.trajectory.x, b.trajectory.x)
        np.testing.assert_array_equal(a.trajectory.t, b.trajectory.t)
        assert a.aperture_hit == b.aperture_hit and a.alive == b.alive
    if packed:                                      # packed blocks come back in global molecule order
        assert [m.trajectory.x[0, 0] for m in merged] == [1000.0 * k for k in range(14)]


def test_identically_seeded_ranks_draw_disjoint_pieces():
    """Host-drawn (custom) distributions under torch.distributed: ranks that seed NumPy's global generator the same
    way must not simulate the same molecules.  Each rank owns a contiguous block of the reference's draw loops
    (trajectory_simulator.py:52-58) and skips the draws of lower ranks, so the pieces, in rank order, are exactly
    the single-process sample."""
    from trajectories import _engine as eng
    from trajectories.distributions import CeNTREXVelocityDistribution, GaussianPositionDistribution

    class Shifted(GaussianPositionDistribution):     # a user-defined Distribution: no device generator for it
        def draw(self, n):
            return super().draw(n) + 1.0

    assert eng.make_source(CeNTREXVelocityDistribution(), Shifted()) is None
    N, loops = 7, 10
    np.random.seed(5)
    whole = list(eng.owned_draws(CeNTREXVelocityDistribution(), Shifted(), N, loops, 0, 1))
    assert len(whole) == loops
    pieces = []
    for rank in range(3):
        np.random.seed(5)                            # every rank runs the same script
        pieces += list(eng.owned_draws(CeNTREXVelocityDistribution(), Shifted(), N, loops, rank, 3))
    assert len(pieces) == loops
    for (v0, x0), (v1, x1) in zip(whole, pieces):
        np.testing.assert_array_equal(v0, v1)
        np.testing.assert_array_equal(x0, x1)
    firsts = {p[0][0, 0] for p in pieces}
    assert len(firsts) == loops                      # no chunk twice
    assert list(eng.owned_draws(CeNTREXVelocityDistribution(), Shifted(), 0, loops, 0, 1)) == []


def _merge_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import json

    import torch.distributed as dist

    from trajectories import _hybrid

    dist.init_process_group("gloo", rank=rank, world_size=world)
    # a user element is free to invent fate names: the ranks need not have met the same ones
    mine = {"4K shield": 10 + rank, "Detected": 3} if rank == 0 else {"4K shield": 10 + rank, "scattered away": 7}
    merged = _hybrid.merge_counts_across_ranks(dict(mine))
    (Path(out_dir) / f"merged{rank}.json").write_text(json.dumps(merged, sort_keys=True))
    dist.destroy_process_group()


def test_hybrid_counter_merge_world2(tmp_path):
    """Mixed beamlines (user-defined elements) keep their Counter as name -> count on the host; under
    torch.distributed the dictionaries of all ranks are summed by name on every rank."""
    import json

    import torch.multiprocessing as mp

    from trajectories import _hybrid

    assert _hybrid.merge_counts_across_ranks({"a": 1}) == {"a": 1}          # no process group: unchanged
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_merge_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    want = {"4K shield": 21, "Detected": 3, "scattered away": 7}
    for rank in (0, 1):
        assert json.loads((tmp_path / f"merged{rank}.json").read_text()) == want
