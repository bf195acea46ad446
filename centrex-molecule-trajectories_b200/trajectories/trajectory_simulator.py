"""TrajectorySimulator / Counter / SimulationResult (reference trajectory_simulator.py:22-254).

`run_simulation` keeps the reference's signature and observable behaviour
(run size 100*n_jobs*int(N_traj/(100*n_jobs)), Counter keys only for fates that
occurred, saved molecules for the apertures of interest, `.counter`, `.result`,
`.results[run_name]`), but the per-molecule Python loop (lines 52-78) and the
joblib process pool (81-83) are replaced by CUDA launches:

  * built-in distributions are sampled on the GPU (Philox indexed by the global
    molecule number), so a run is a stream of kernel launches with no host data;
  * any other Distribution is drawn on the host exactly as the reference does
    (`vdist.draw(N)` then `xdist.draw(N)` per chunk) and replayed on the GPU;
  * molecules whose fate is of interest are re-propagated by the trajectory
    kernel and come back as Molecule objects with full (n,3) x/v/a and (n,) t;
  * under torch.distributed (one process per GPU) each rank takes a contiguous
    block of the global index range and the per-fate counts are all-reduced.

`n_jobs` only enters through the reference's run-size arithmetic.
"""
from __future__ import annotations

import os
from copy import copy
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional

import numpy as np

from . import _engine as eng
from . import _hybrid
from .beamline import Beamline
from .distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution, Distribution
from .molecule import Molecule, SavedMolecules

__all__ = ["TrajectorySimulator", "Counter", "SimulationResult"]


class Counter:
    """Per-fate molecule counts (trajectory_simulator.py:106-177)."""

    def __init__(self) -> None:
        self.counter_dict = {}

    def increment_counter(self, aperture_hit: str, by: int = 1) -> None:
        self.counter_dict[aperture_hit] = self.counter_dict.get(aperture_hit, 0) + by

    def print(self) -> None:
        print("Number of molecules that hit each element:")
        for key, value in self.counter_dict.items():
            print(f"{key} : {value}")

    def calculate_efficiency(self) -> float:
        total = sum(self.counter_dict.values())
        if "Detected" in self.counter_dict:
            return self.counter_dict["Detected"] / total
        return 0

    def merge_counters(self, others: List["Counter"]) -> None:
        for other in others:
            for key, value in other.counter_dict.items():
                self.increment_counter(key, value)

    def save_to_hdf(self, filepath: Path, run_name: str) -> None:
        from ._hdf import save_counter

        save_counter(self, filepath, run_name)


@dataclass
class SimulationResult:
    """counter, beamline, xdist, vdist, molecules: the reference's five fields (trajectory_simulator.py:180-191).

    Under torch.distributed every rank holds the molecules of its own block of the global index range, in global
    order; `molecule_offset` is the number of molecules saved by lower ranks (from an all_gather of the per-rank
    counts) and `n_molecules_total` the sum over ranks, so that `save_to_hdf` numbers `molecule_<i>` globally and
    the file written by N ranks equals the file of a single-GPU run.
    """
    counter: Counter
    beamline: Beamline
    xdist: Distribution
    vdist: Distribution
    molecules: List[Molecule]
    molecule_offset: int = field(default=0, compare=False)
    n_molecules_total: Optional[int] = field(default=None, compare=False)

    def plot(self, N_max: int = 10000, elements: List[str] = None, show: bool = True):
        axes = self.beamline.plot()
        shown = 0
        for molecule in self.molecules:
            if shown >= N_max:
                break
            if elements is not None and molecule.aperture_hit not in elements:
                continue
            molecule.plot_trajectory(axes)
            shown += 1
        if show:
            import matplotlib.pyplot as plt

            plt.show()
        return axes

    def save_to_hdf(self, filepath: Path, run_name: str, packed: bool = False) -> None:
        """The reference's layout (trajectory_simulator.py:218-254).  `packed=True` stores the trajectories as one
        `(total_rows, 10)` dataset with row offsets instead of a group and four datasets per molecule.

        Under torch.distributed rank 0 writes the run's metadata, then the ranks append their molecules one after
        the other in rank order (the file must be on a filesystem all ranks see)."""
        from ._hdf import h5py

        rank, world = eng.dist_info()
        proceed = True
        if rank == 0:
            with h5py().File(filepath, "a") as f:
                try:
                    f.create_group(run_name)
                except ValueError:
                    if input("Run name already exists. Overwrite? y/n") != "y":
                        proceed = False
                    else:
                        del f[run_name]
                        f.create_group(run_name)
            if proceed:
                for key in ("counter", "beamline", "xdist", "vdist"):
                    getattr(self, key).save_to_hdf(filepath, run_name)
        proceed = eng.broadcast_object(proceed)
        if not proceed:
            return
        for turn in range(world):
            if turn == rank:
                self.save_molecules_to_hdf(filepath, run_name, packed=packed)
            eng.barrier()

    def save_molecules_to_hdf(self, filepath: Path, run_name: str, packed: bool = False) -> None:
        from ._hdf import h5py

        print("Saving trajectories...")
        with h5py().File(filepath, "a") as f:
            if packed:
                self._save_packed(f, run_name)
                return
            for i, molecule in enumerate(self.molecules, self.molecule_offset):
                molecule.save_to_hdf(f, run_name, f"trajectories/molecule_{i}")

    def _save_packed(self, f, run_name: str) -> None:
        """<run>/trajectories_packed[/rank_<r>]: rows (total_rows, 10) = x,y,z,vx,vy,vz,ax,ay,az,t; offsets (n+1,);
        fate (n,) indices into the attribute `fate_names`; alive (n,); attribute first_molecule = global number of
        molecule 0 of the block."""
        rank, world = eng.dist_info()
        path = run_name + "/trajectories_packed" + (f"/rank_{rank}" if world > 1 else "")
        g = f.create_group(path)
        names = sorted({m.aperture_hit for m in self.molecules})
        counts = np.array([m.trajectory.n for m in self.molecules], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        rows = np.empty((int(offsets[-1]), 10))
        for k, m in enumerate(self.molecules):
            tr = m.trajectory
            lo, hi = offsets[k], offsets[k + 1]
            rows[lo:hi, 0:3], rows[lo:hi, 3:6], rows[lo:hi, 6:9], rows[lo:hi, 9] = tr.x, tr.v, tr.a, tr.t
        g.create_dataset("rows", data=rows)
        g.create_dataset("offsets", data=offsets)
        g.create_dataset("fate", data=np.array([names.index(m.aperture_hit) for m in self.molecules], dtype=np.int32))
        g.create_dataset("alive", data=np.array([bool(m.alive) for m in self.molecules], dtype=np.uint8))
        g.attrs["fate_names"] = np.array(names, dtype=object) if names else np.array([], dtype="S1")
        g.attrs["first_molecule"] = int(self.molecule_offset)


class TrajectorySimulator:
    """Runs trajectory simulations on the GPU and stores the results."""

    def __init__(self, device=None, seed: Optional[int] = None, chunk: int = eng.DEFAULT_CHUNK,
                 math: str = "exact") -> None:
        """math="exact" (default): every operation rounds as in the reference, in the reference's order -- fates are
        the reference's fates and rows agree to the last bit except where DESIGN.md "Exact arithmetic" lists a known
        difference (NumPy's scalar `dt**2` goes through libm pow(): <= 1 ulp in y in ~1e-5 of ballistic steps; a
        literal -0.0 input becomes +0.0).  math="contracted" runs the same algorithm with fused multiply-adds and
        reciprocal multiplications (agreement to ~1e-13 relative, 1e-9 in the worst case on coordinates that pass
        near zero; 1.2x the lens-integrator throughput)."""
        self.counter = Counter()
        self.results = {}
        self.device = device
        self.math = math
        self.seed = seed
        self.chunk = int(chunk)
        self.last_work = None   # [ballistic rows, lens RK steps, table out-of-range evals, lens entries]

    # -- public API -----------------------------------------------------------
    def run_simulation(
        self,
        beamline: Beamline,
        run_name: str,
        vdist=CeNTREXVelocityDistribution(),
        xdist=CeNTREXPositionDistribution(),
        N_traj: int = 1000,
        apertures_of_interest=[],
        n_jobs=1,
        seed: Optional[int] = None,
    ) -> None:
        torch = eng._torch()
        # run size exactly as trajectory_simulator.py:48-49 (the remainder is dropped)
        N_loops = 100 * n_jobs
        N = int(N_traj / N_loops)
        total = N * N_loops

        if _hybrid.is_hybrid(beamline.elements):
            # user-defined elements: the built-in runs on the GPU, the user's propagate_through on the host (_hybrid.py)
            return self._run_hybrid(beamline, run_name, vdist, xdist, N, N_loops, total, apertures_of_interest, seed)

        flat = eng.flatten(beamline.elements)
        prop = eng.Propagator(flat, self.device, math=self.math)
        prop.reset()
        save_mask = flat.save_mask(list(apertures_of_interest))
        rank, world = eng.dist_info()
        source = eng.make_source(vdist, xdist)
        molecules = SavedMolecules()

        if source is not None:
            seed = self._pick_seed(seed, prop.tdev)
            lo, hi = eng.shard_range(total, rank, world)
            lo0 = lo
            if hi - lo > 8 * eng.PILOT_MOLECULES:
                # A large run starts with a pilot launch whose lens queue could hold every molecule; the fraction that
                # actually reaches the lens (0.54 % for the CeNTREX source) then sizes the queues of all later launches:
                # 128 B x n per stream slot would be 8.6 GB at the default chunk, to hold a queue that is 0.5 % full.
                # The result does not depend on where a run is cut.  A queue that overflows all the same (work[6]) sends
                # the whole run through full-size queues below.
                res = prop.propagate_philox(source, seed, lo, eng.PILOT_MOLECULES, save_mask=save_mask)
                if save_mask and res.saved_index.numel():
                    molecules.extend(self._collect(prop, prop.draw(source, seed, index=res.saved_index)))
                prop.learn_entry_fraction(eng.PILOT_MOLECULES)
                lo += eng.PILOT_MOLECULES
            chunk = prop.fit_chunk(self.chunk)
            if not save_mask and (1 << 21) <= hi - lo <= chunk:
                # a run that fits one chunk goes out as two halves on two streams: the walk kernel of the
                # second overlaps the lens segments of the first (the result does not depend on the cut)
                chunk = (hi - lo + 1) // 2
            for k, first in enumerate(range(lo, hi, chunk)):
                n = min(chunk, hi - first)
                if not save_mask:
                    # nothing to read back per chunk: alternate streams so consecutive chunks overlap
                    prop.propagate_philox(source, seed, first, n, slot=k)
                    continue
                res = prop.propagate_philox(source, seed, first, n, save_mask=save_mask)
                if res.saved_index.numel():
                    ic = prop.draw(source, seed, index=res.saved_index)
                    molecules.extend(self._collect(prop, ic))
            prop.join()
            if prop.entry_fraction is not None and prop.queue_overflow():
                # the sized queues were too small after all: once more, with queues that hold everything
                prop.reset()
                prop.entry_fraction = None
                molecules = SavedMolecules()
                chunk = prop.fit_chunk(self.chunk)
                for k, first in enumerate(range(lo0, hi, chunk)):
                    n = min(chunk, hi - first)
                    res = prop.propagate_philox(source, seed, first, n, save_mask=save_mask, slot=None if save_mask else k)
                    if save_mask and res.saved_index.numel():
                        molecules.extend(self._collect(prop, prop.draw(source, seed, index=res.saved_index)))
                prop.join()
        else:
            # Host draws, replayed.  The reference's N_loops chunks are split into one contiguous block of loops per
            # rank.  A Distribution object is opaque (it may draw from NumPy's global RNG, which scripts seed the same
            # way in every process), so every rank walks through the draws in the reference's order and throws away
            # the chunks of lower ranks: identically seeded ranks then simulate DISJOINT pieces of one and the same
            # sample -- the union, in rank order, is the single-process run -- instead of world copies of the same
            # molecules; differently seeded ranks simulate independent samples.  Costs rank r its share of the
            # draws times r + 1; the built-in distributions never come here (Philox on the device).
            batch_v, batch_x, filled = [], [], 0

            def flush():
                nonlocal batch_v, batch_x, filled
                if not filled:
                    return
                host = np.concatenate([np.concatenate(batch_x, axis=1), np.concatenate(batch_v, axis=1)], axis=0)
                ic = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float64)).to(prop.tdev)
                res = prop.propagate_ic(ic, want_fate=False, save_mask=save_mask)
                if save_mask and res.saved_index.numel():
                    molecules.extend(self._collect(prop, ic, select=res.saved_index))
                batch_v, batch_x, filled = [], [], 0

            for vs, xs in eng.owned_draws(vdist, xdist, N, N_loops, rank, world):
                batch_v.append(vs)
                batch_x.append(xs)
                filled += N
                if filled >= self.chunk:
                    flush()
            flush()

        counts = eng.allreduce_counts(prop.counters.clone()).cpu().numpy()
        work = eng.allreduce_counts(prop.work.clone()).cpu().numpy()
        self.last_work = work
        prop.release()          # the queue workspaces go back to the allocator; the beamline handle stays cached
        if work[2] > 0:
            # the reference's interp1d raises for r beyond the table (bounds_error=True)
            raise ValueError(
                f"A value in x_new is above the interpolation range ({int(work[2])} lens force "
                "evaluations fell outside the a_interp table)")

        self.counter = Counter()
        for name, c in zip(flat.fate_names, counts):
            if c > 0:
                self.counter.increment_counter(name, int(c))
        # the gather of the saved-trajectory counts: every rank keeps its own molecules (rank order = global order) and
        # learns the global number of its first one, so that save_to_hdf numbers molecule_<i> as a single-GPU run does
        saved_counts = eng.gather_counts(len(molecules), prop.device)
        offset, n_saved = sum(saved_counts[:rank]), sum(saved_counts)
        self.result = SimulationResult(self.counter, beamline, xdist, vdist, molecules, offset, n_saved)
        self.results[run_name] = SimulationResult(self.counter, beamline, xdist, vdist, molecules, offset, n_saved)

    def _pick_seed(self, seed, tdev) -> int:
        """Seed of the device source: the call's, the simulator's, or one drawn from NumPy's global RNG (which honours
        np.random.seed() as the reference's draws do) and agreed on by all ranks."""
        torch = eng._torch()
        if seed is None:
            seed = self.seed
        if seed is None:
            seed = int(np.random.randint(0, 2**62))
            _rank, world = eng.dist_info()
            if world > 1:
                s = torch.tensor([seed], dtype=torch.int64, device=tdev if torch.distributed.get_backend() == "nccl" else "cpu")
                torch.distributed.broadcast(s, 0)
                seed = int(s.item())
        return int(seed)

    def _run_hybrid(self, beamline, run_name, vdist, xdist, N, N_loops, total, apertures_of_interest, seed) -> None:
        """run_simulation for a beamline with user-defined elements: same sample, same sharding, same result objects;
        the Counter is keyed by whatever fate names the user's elements report."""
        torch = eng._torch()
        run = _hybrid.HybridRun(beamline.elements, self.device, self.math, list(apertures_of_interest))
        rank, world = eng.dist_info()
        source = eng.make_source(vdist, xdist)
        if source is not None:
            seed = self._pick_seed(seed, run.tdev)
            lo, hi = eng.shard_range(total, rank, world)
            step = min(self.chunk, _hybrid.HYBRID_CHUNK)
            for first in range(lo, hi, step):
                run.run_chunk(_hybrid.draw_ic(source, seed, first, min(step, hi - first), run.device))
        else:
            for vs, xs in eng.owned_draws(vdist, xdist, N, N_loops, rank, world):
                host = np.ascontiguousarray(np.concatenate([xs, vs], axis=0), dtype=np.float64)
                run.run_chunk(torch.from_numpy(host).to(run.tdev))
        self.last_work = run.work
        if run.work[2] > 0:
            raise ValueError(f"A value in x_new is above the interpolation range ({int(run.work[2])} lens force "
                             "evaluations fell outside the a_interp table)")
        self.counter = Counter()
        for name, c in _hybrid.merge_counts_across_ranks(run.counts).items():
            self.counter.increment_counter(name, int(c))
        saved_counts = eng.gather_counts(len(run.molecules), run.device)
        offset, n_saved = sum(saved_counts[:rank]), sum(saved_counts)
        self.result = SimulationResult(self.counter, beamline, xdist, vdist, run.molecules, offset, n_saved)
        self.results[run_name] = SimulationResult(self.counter, beamline, xdist, vdist, run.molecules, offset, n_saved)

    # the reference's README calls the parallel entry point by this name (README.md:52)
    run_simulation_parallel = run_simulation

    def run_sweep(
        self,
        beamline: Beamline,
        states,
        voltages=None,
        run_name: str = "J = {J}, mJ = {mJ}, V = {V:.0f}",
        vdist=CeNTREXVelocityDistribution(),
        xdist=CeNTREXPositionDistribution(),
        N_traj: int = 1000,
        apertures_of_interest=[],
        n_jobs=1,
        seed: Optional[int] = None,
        lens_name: str = "ES lens",
    ) -> dict:
        """The loop of examples/lens_simulation_different_states.py:135-152 (set the lens' state, reset its table,
        run, keep the result) over `states` x `voltages` as ONE batched call.

        Every point is what `run_simulation(beamline, name, vdist, xdist, N_traj, apertures_of_interest, n_jobs, seed)`
        gives after `lens.state = state; lens.V = V; lens.a_interp = None` -- same Counter, same saved molecules --
        but the Stark tables of all points are built in one vectorised pass, the beamline is flattened once, and,
        when no trajectories are asked for, the launches of consecutive points go out on alternating streams with a
        single read-back of all Counters at the end.  All points use the same seed, i.e. the same molecules (common
        random numbers: differences between points are differences of the lens, not of the sample).

        `states`: objects with `find_largest_component()` (centrex_TlF states) or `(J, mJ)` pairs; `voltages`: None
        keeps the lens' own voltage.  Returns `{(J, mJ, V): SimulationResult}`, also stored in `self.results` under
        `run_name.format(J=, mJ=, V=)`; `self.counter` / `self.result` are those of the last point.
        """
        from copy import copy as shallow
        from .stark_potential import state_quantum_numbers

        torch = eng._torch()
        lens = beamline.find_element(lens_name)
        if lens is None:
            raise ValueError(f"no element named {lens_name!r} in the beamline")
        source = eng.make_source(vdist, xdist)
        if source is None:
            raise ValueError("run_sweep needs the built-in distributions (device source); loop over run_simulation otherwise")
        Vs = [lens.V] if voltages is None else [float(v) for v in voltages]
        points = [(state, V) for state in states for V in Vs]
        N_loops = 100 * n_jobs
        total = int(N_traj / N_loops) * N_loops
        rank, world = eng.dist_info()
        if seed is None:
            seed = self.seed
        if seed is None:
            seed = int(eng.broadcast_object(int(np.random.randint(0, 2**62))))

        # the Stark tables of all points, built side by side (each is a chain of small LAPACK calls that release
        # the GIL; points whose states share an mF block and a voltage share one set of eigenpairs, _tlf_full.py),
        # then one flattening per point, differing only in the lens table
        def table_of(point):
            probe = shallow(lens)
            probe.state, probe.V, probe.a_interp = point[0], point[1], None
            return probe.ensure_a_interp()

        from concurrent.futures import ThreadPoolExecutor
        from contextlib import nullcontext

        try:        # the matrices are 26 x 26 at most: BLAS threads of their own only get in each other's way
            from threadpoolctl import threadpool_limits
            one_blas_thread = threadpool_limits(1)
        except Exception:
            one_blas_thread = nullcontext()
        with one_blas_thread, ThreadPoolExecutor(max_workers=max(1, min(len(points), os.cpu_count() or 1, 16))) as pool:
            interps = list(pool.map(table_of, points))
        flats, keys, lenses = [], [], []
        saved_state, saved_V, saved_tab = lens.state, lens.V, lens.a_interp
        try:
            for (state, V), interp in zip(points, interps):
                lens.state, lens.V, lens.a_interp = state, V, interp
                flats.append(eng.flatten(beamline.elements))
                J, mJ = state_quantum_numbers(state)
                keys.append((J, mJ, V))
                lenses.append((state, V, lens.a_interp))
        finally:
            lens.state, lens.V, lens.a_interp = saved_state, saved_V, saved_tab

        prop = eng.Propagator(flats[0], self.device, math=self.math)
        save_mask = flats[0].save_mask(list(apertures_of_interest))
        lo, hi = eng.shard_range(total, rank, world)
        chunk = prop.fit_chunk(self.chunk)
        if not save_mask and (1 << 21) <= hi - lo <= chunk:
            chunk = (hi - lo + 1) // 2
        # How many molecules reach the lens does not depend on the lens table, so a small pilot launch of the first
        # point sizes the lens queues of every launch of the sweep (128 B x n per stream slot otherwise: 1.9 GB of
        # fresh device memory for 1e7-molecule points, 0.16 s of a first call); a queue that overflows all the same
        # (work[6]) sends the whole sweep through full-size queues.
        if hi - lo > 4 * eng.SWEEP_PILOT_MOLECULES:
            prop.propagate_philox(source, seed, lo, eng.SWEEP_PILOT_MOLECULES)
            prop.learn_entry_fraction(eng.SWEEP_PILOT_MOLECULES)

        def launch_all():
            per_point, k = [], 0
            for flat in flats:
                prop.rebind(flat)
                molecules = SavedMolecules()
                for first in range(lo, hi, chunk):
                    n = min(chunk, hi - first)
                    if not save_mask:
                        prop.propagate_philox(source, seed, first, n, slot=k)
                        k += 1
                        continue
                    res = prop.propagate_philox(source, seed, first, n, save_mask=save_mask)
                    if res.saved_index.numel():
                        molecules.extend(self._collect(prop, prop.draw(source, seed, index=res.saved_index)))
                per_point.append((prop.counters, prop.work, molecules))
            prop.join()
            return per_point

        per_point = launch_all()
        works = torch.stack([w for _, w, _ in per_point]).cpu().numpy()
        if prop.entry_fraction is not None and works[:, 6].sum() > 0:
            prop.entry_fraction = None
            per_point = launch_all()
        counts = eng.allreduce_counts(torch.stack([c for c, _, _ in per_point])).cpu().numpy()
        works = eng.allreduce_counts(torch.stack([w for _, w, _ in per_point])).cpu().numpy()
        self.last_work = works.sum(axis=0)
        if works[:, 2].sum() > 0:
            raise ValueError(f"A value in x_new is above the interpolation range ({int(works[:, 2].sum())} lens force "
                             "evaluations fell outside the a_interp table)")
        out = {}
        for p, (key, flat) in enumerate(zip(keys, flats)):
            counter = Counter()
            for name, c in zip(flat.fate_names, counts[p]):
                if c > 0:
                    counter.increment_counter(name, int(c))
            # each result carries its own beamline, with the lens as it was for that point
            bl = shallow(beamline)
            bl.elements = [shallow(e) if e is lens else e for e in beamline.elements]
            pl = bl.elements[beamline.elements.index(lens)]
            pl.state, pl.V, pl.a_interp = lenses[p]
            molecules = per_point[p][2]
            saved_counts = eng.gather_counts(len(molecules), prop.device) if world > 1 else [len(molecules)]
            res = SimulationResult(counter, bl, xdist, vdist, molecules, sum(saved_counts[:rank]), sum(saved_counts))
            out[key] = res
            self.results[run_name.format(J=key[0], mJ=key[1], V=key[2])] = res
            self.counter, self.result = counter, res
        prop.release()
        return out

    def plane_distributions(
        self,
        beamline: Beamline,
        z,
        elements: Optional[List[str]] = None,
        vdist=CeNTREXVelocityDistribution(),
        xdist=CeNTREXPositionDistribution(),
        N_traj: int = 1000,
        n_jobs=1,
        seed: Optional[int] = None,
    ):
        """Positions and velocities at the plane(s) `z` without storing any trajectory.

        Same molecules and same numbers as `run_simulation(..., apertures_of_interest=elements)` followed by
        `post_processing.find_radial_pos_dist(result, z, elements)` / `find_vel_dist` (reference
        post_processing.py:20-140), but the crossing is evaluated inside the propagation kernel, so memory
        is 40 B per molecule and plane instead of up to 49 kB per saved trajectory.  `elements=None` keeps
        every molecule that reaches the plane.  Returns `(xy [m, 2], v [m, 3])` for a scalar `z`, a list
        of such pairs for a sequence; molecules in global-index order (this rank's block under
        torch.distributed).  Also updates `.counter` like a run would.
        """
        torch = eng._torch()
        scalar = np.ndim(z) == 0
        zs = np.atleast_1d(np.asarray(z, dtype=np.float64))
        N_loops = 100 * n_jobs
        N = int(N_traj / N_loops)
        total = N * N_loops
        flat = eng.flatten(beamline.elements)
        prop = eng.Propagator(flat, self.device, math=self.math)
        prop.reset()
        mask = flat.save_mask(list(elements)) if elements is not None else None
        rank, world = eng.dist_info()
        source = eng.make_source(vdist, xdist)
        chunk = min(self.chunk, 1 << 24)
        xy = [[] for _ in zs]
        vel = [[] for _ in zs]

        def probe(ic, select=None):
            out, valid, fate = prop.plane_crossings(ic, zs, select=select)
            if mask is None:     # every molecule goes through the crossing kernel: its fates are the Counter
                prop.counters[: len(flat.fate_names)] += torch.bincount(fate.long(), minlength=len(flat.fate_names))
            for p in range(len(zs)):
                cols = out[p][:, valid[p]].cpu().numpy()
                xy[p].append(cols[0:2].T)
                vel[p].append(cols[2:5].T)

        if source is not None:
            if seed is None:
                seed = self.seed
            if seed is None:
                seed = int(np.random.randint(0, 2**62))
                if world > 1:
                    s = torch.tensor([seed], dtype=torch.int64, device=prop.tdev if torch.distributed.get_backend() == "nccl" else "cpu")
                    torch.distributed.broadcast(s, 0)
                    seed = int(s.item())
            lo, hi = eng.shard_range(total, rank, world)
            for first in range(lo, hi, chunk):
                n = min(chunk, hi - first)
                if mask is None:
                    probe(prop.draw(source, seed, first, n))
                elif mask:
                    res = prop.propagate_philox(source, seed, first, n, save_mask=mask)
                    if res.saved_index.numel():
                        probe(prop.draw(source, seed, index=res.saved_index))
                else:
                    prop.propagate_philox(source, seed, first, n)
        else:
            for vs, xs in eng.owned_draws(vdist, xdist, N, N_loops, rank, world):
                ic = torch.from_numpy(np.ascontiguousarray(np.concatenate([xs, vs]), dtype=np.float64)).to(prop.tdev)
                if mask is None:
                    probe(ic)
                elif mask:
                    res = prop.propagate_ic(ic, want_fate=False, save_mask=mask)
                    if res.saved_index.numel():
                        probe(ic, select=res.saved_index)
                else:
                    prop.propagate_ic(ic, want_fate=False)
        prop.join()

        counts = eng.allreduce_counts(prop.counters.clone()).cpu().numpy()
        self.counter = Counter()
        for name, c in zip(flat.fate_names, counts):
            if c > 0:
                self.counter.increment_counter(name, int(c))
        pairs = [(np.concatenate(a) if a else np.empty((0, 2)), np.concatenate(b) if b else np.empty((0, 3)))
                 for a, b in zip(xy, vel)]
        return pairs[0] if scalar else pairs

    # -- helpers ----------------------------------------------------------------
    @staticmethod
    def _collect(prop: "eng.Propagator", ic, select=None) -> SavedMolecules:
        rows, offsets, fate, rows_arrived = prop.trajectories(ic, select=select, defer_sync=True)
        lo = np.asarray(offsets).tolist()
        rows_arrived()
        # (a non-finite value poisons every later row of its trajectory, so the last rows tell)
        last = rows[np.asarray(lo[1:], dtype=np.int64) - 1] if len(lo) > 1 else rows[:0]
        strip = bool(last.size) and not np.isfinite(last).all()
        # each trajectory is a view of its slice of the result block, wrapped in a Molecule when it is first asked for
        mols = SavedMolecules()
        mols.add_rows(rows, lo, np.asarray(fate).tolist(), prop.flat.fate_names, strip)
        return mols
