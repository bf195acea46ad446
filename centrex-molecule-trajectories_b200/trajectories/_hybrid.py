"""Beamlines that mix built-in elements (CUDA) with user-defined `BeamlineElement` subclasses (Python).

The reference's plugin API is the abstract `BeamlineElement` (apertures.py:22-54): anything with `z0`, `L`,
`N_steps()` and `propagate_through(molecule)` can be put into a `Beamline`, and `Beamline.propagate_through`
(beamline.py:20-38) simply calls the elements in flight order until the molecule is dead.  The CUDA path knows the
five element types of the reference; an element it does not know -- a subclass written by the user, or a subclass
of a built-in type that overrides `propagate_through` -- is executed on the host, by the user's own code, on exactly
the molecules that reach it:

    [built-in elements]  ->  user element  ->  [built-in elements]  ->  ...
     one device handle       Python, per        next device handle
     cmt_propagate_ic        molecule (or       cmt_resume (from the
     fate + last row         vectorised)        row the element left)

* a run of built-in elements is one `cmt_beamline_t`; its "Detected" fate means "left this run alive";
* the survivors' last rows (x, v, a, t -- 80 B each) come to the host, every one becomes a `Molecule` whose
  trajectory holds that row, and `element.propagate_through(molecule)` runs unchanged (it sees `molecule.x()`,
  `.v()`, `.a()`, `.t()`, `update_trajectory`, `set_dead`, `set_aperture_hit` as in the reference);
* the molecules it leaves alive go back to the device and resume from their new last row (`cmt_resume`).

An element may additionally offer `propagate_through_batch(rows) -> (alive, rows_out, names)` (an extension of this
build, not part of the reference API): `rows` is a float64 array [k, 10] = x,y,z,vx,vy,vz,ax,ay,az,t of the molecules
that reach it; it returns a bool array [k], the rows the molecules end on [k, 10] and, for the dead ones, the fate
name (one string or a sequence of k strings).  It replaces the per-molecule Python loop; trajectories of such an
element have one row per molecule (the row it returns).

Saved trajectories (`apertures_of_interest`) are assembled from the rows the device segments produce
(`cmt_trajectories`, deterministic, recomputed only for the molecules of interest) and the rows the user's element
appended during the one and only time it was called for that molecule -- a stochastic element is never re-run.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _engine as eng
from . import _native as nat
from .molecule import Molecule, Trajectory

HYBRID_CHUNK = 1 << 22      # molecules per pass: 80 B of last row + 1 B of fate each on the device


def _builtin_types():
    from .beamline_elements.apertures import CircularAperture, FieldPlates, RectangularAperture
    from .beamline_elements.electrostatic_lens import ElectrostaticLens
    from .beamline_elements.meshes import Honeycomb

    return (CircularAperture, RectangularAperture, FieldPlates, ElectrostaticLens, Honeycomb)


def runs_on_device(element) -> bool:
    """True for an element of one of the five built-in types that has not replaced `propagate_through`."""
    for base in _builtin_types():
        if isinstance(element, base):
            return type(element).propagate_through is base.propagate_through
    return False


def is_hybrid(elements: Sequence) -> bool:
    return any(not runs_on_device(e) for e in elements)


class _DeviceStage:
    def __init__(self, elements, device, math):
        self.elements = list(elements)
        self.prop = eng.Propagator(eng.flatten(self.elements), device, math=math)
        self.names = self.prop.flat.fate_names
        self.alive_id = self.prop.flat.fate_detected


class _HostStage:
    def __init__(self, element):
        self.element = element
        self.batch = getattr(element, "propagate_through_batch", None)


def split_stages(elements: Sequence, device, math) -> list:
    """Flight-ordered elements -> alternating device / host stages."""
    stages, run = [], []
    for e in elements:
        if runs_on_device(e):
            run.append(e)
            continue
        if run:
            stages.append(_DeviceStage(run, device, math))
            run = []
        for attr in ("propagate_through", "z0"):
            if not hasattr(e, attr):
                raise TypeError(f"beamline element {type(e).__name__!r} has no {attr}")
        stages.append(_HostStage(e))
    if run:
        stages.append(_DeviceStage(run, device, math))
    return stages


def _resume(prop: eng.Propagator, state):
    """state: device float64 [10, k] -> (fate uint8 [k], last row [10, k]) through cmt_resume."""
    torch = eng._torch()
    k = state.shape[1]
    fate = torch.empty(k, dtype=torch.uint8, device=prop.tdev)
    last = torch.empty((10, k), dtype=torch.float64, device=prop.tdev)
    if k:
        assert state.dtype == torch.float64 and state.shape[0] == 10 and state.stride(1) == 1
        with torch.cuda.device(prop.device):
            nat.check(nat.lib().cmt_resume(prop.dev.handle, k, state.data_ptr(), state.stride(0), last.data_ptr(), k,
                                           None, fate.data_ptr(), eng._stream_ptr(prop.device)))
    return fate, last


def _initial_rows(ic: np.ndarray) -> np.ndarray:
    """[6, k] initial conditions -> rows [k, 10] as Molecule.init_trajectory stores them (molecule.py:15-24)."""
    k = ic.shape[1]
    rows = np.zeros((k, 10))
    rows[:, 0:6] = ic.T
    rows[:, 7] = -eng.G
    return rows


def _molecule_at(row: np.ndarray, element) -> Molecule:
    try:
        spare = int(element.N_steps())
    except Exception:
        spare = 0
    mol = Molecule()
    mol.trajectory = Trajectory(n_rows=1 + max(spare, 0) + 10)
    mol.trajectory.update(row[0:3], row[3:6], row[6:9], row[9])
    return mol


class HybridRun:
    """One run over a mixed beamline: Counter by fate name, optionally the saved molecules."""

    def __init__(self, elements: Sequence, device, math: str, apertures_of_interest: Sequence[str]):
        self.device = eng.resolve_device(device)
        self.stages = split_stages(elements, self.device, math)
        self.interest = set(apertures_of_interest)
        self.counts: Dict[str, int] = {}
        self.molecules: List[Molecule] = []
        self.work = np.zeros(8, dtype=np.int64)
        self.tdev = eng._torch().device("cuda", self.device)

    # ------------------------------------------------------------------------------------------------
    def _count(self, name: str, k: int) -> None:
        if k:
            self.counts[name] = self.counts.get(name, 0) + int(k)

    def run_chunk(self, ic) -> None:
        """ic: device float64 [6, n] -- the chunk's initial conditions."""
        torch = eng._torch()
        n = ic.shape[1]
        if n == 0:
            return
        saving = bool(self.interest)
        alive = torch.arange(n, device=self.tdev)            # chunk-local indices of the molecules still in flight
        state = None                                           # device [10, k]; None = initial conditions
        # for the saved trajectories: where every molecule ended, and what the host elements appended
        end_stage = np.full(n, len(self.stages), dtype=np.int32) if saving else None
        end_name: Dict[int, str] = {}
        fragments: List[Dict[int, np.ndarray]] = [dict() for _ in self.stages]

        for si, stage in enumerate(self.stages):
            if alive.numel() == 0:
                break
            if isinstance(stage, _DeviceStage):
                last = None
                if state is None:
                    # the normal launch, fates only (so the FP32 filter of the walk kernel stays on); the last rows of
                    # the few survivors follow below from a resume launch on just those
                    stage.prop.reset()
                    res = stage.prop.propagate_ic(ic, want_fate=True)
                    fate = res.fate
                    self.work += res.work.cpu().numpy()
                else:
                    fate, last = _resume(stage.prop, state)
                hist = torch.bincount(fate.long(), minlength=len(stage.names)).cpu().numpy()
                for fid, name in enumerate(stage.names):
                    if fid != stage.alive_id:
                        self._count(name, hist[fid])
                going = fate == stage.alive_id
                if saving:
                    stopped = (~going).nonzero().squeeze(1)
                    if stopped.numel():
                        idx = alive[stopped].cpu().numpy()
                        ids = fate[stopped].cpu().numpy()
                        keep = np.isin(ids, [k for k, nm in enumerate(stage.names) if nm in self.interest])
                        end_stage[idx] = si
                        for j, f in zip(idx[keep], ids[keep]):
                            end_name[int(j)] = stage.names[int(f)]
                        end_stage[idx[~keep]] = -1            # of no interest
                sel = going.nonzero().squeeze(1)
                if last is None:
                    start = torch.zeros((10, sel.numel()), dtype=torch.float64, device=self.tdev)
                    start[0:6] = ic[:, sel]
                    start[7] = -eng.G
                    again, state = _resume(stage.prop, start)
                    if sel.numel() and not bool((again == stage.alive_id).all()):
                        raise RuntimeError("resume launch disagrees with the propagation launch about a survivor")
                else:
                    state = last[:, sel].contiguous()
                alive = alive[sel]
                continue

            # ---- host stage: the user's element, on the molecules that reach it ----
            idx = alive.cpu().numpy()
            rows = _initial_rows(ic[:, alive].cpu().numpy()) if state is None else np.ascontiguousarray(state.cpu().numpy().T)
            keep_alive = np.zeros(len(idx), dtype=bool)
            new_rows = rows.copy()
            if stage.batch is not None:
                ok, out, names = stage.batch(rows.copy())
                ok = np.asarray(ok, dtype=bool)
                out = np.asarray(out, dtype=np.float64)
                if ok.shape != (len(idx),) or out.shape != rows.shape:
                    raise ValueError(f"{type(stage.element).__name__}.propagate_through_batch must return "
                                     f"(bool[{len(idx)}], float64[{len(idx)}, 10], names)")
                keep_alive, new_rows = ok, out
                for j in np.nonzero(~ok)[0]:
                    name = names if isinstance(names, str) else names[j]
                    self._count(name, 1)
                    if saving:
                        end_stage[idx[j]] = si if name in self.interest else -1
                        end_name[int(idx[j])] = name
                if saving:
                    for j in range(len(idx)):
                        fragments[si][int(idx[j])] = out[j:j + 1]
            else:
                for j in range(len(idx)):
                    mol = _molecule_at(rows[j], stage.element)
                    stage.element.propagate_through(mol)
                    tr = mol.trajectory
                    if mol.alive:
                        keep_alive[j] = True
                    else:
                        name = getattr(mol, "aperture_hit", stage.element.name)
                        self._count(name, 1)
                        if saving:
                            end_stage[idx[j]] = si if name in self.interest else -1
                            end_name[int(idx[j])] = name
                    k = tr.n - 1
                    new_rows[j, 0:3], new_rows[j, 3:6], new_rows[j, 6:9], new_rows[j, 9] = tr.x[k], tr.v[k], tr.a[k], tr.t[k]
                    if saving:
                        fragments[si][int(idx[j])] = tr.as_rows()[1:]
            going_rows = new_rows[keep_alive]
            if going_rows.size and np.any(going_rows[:, 8] != 0.0):
                raise ValueError(f"{type(stage.element).__name__} left a molecule with a_z != 0: not supported on the GPU path")
            alive = alive[torch.from_numpy(np.nonzero(keep_alive)[0]).to(self.tdev)]
            state = torch.from_numpy(np.ascontiguousarray(going_rows.T)).to(self.tdev)

        self._count("Detected", alive.numel())
        if saving:
            if "Detected" in self.interest:
                for j in alive.cpu().numpy():
                    end_name[int(j)] = "Detected"
            else:
                end_stage[alive.cpu().numpy()] = -1
            self._assemble(ic, end_stage, end_name, fragments)

    # ------------------------------------------------------------------------------------------------
    def _assemble(self, ic, end_stage, end_name, fragments) -> None:
        """Trajectories of the molecules of interest: device rows recomputed segment by segment, host rows as kept."""
        torch = eng._torch()
        chosen = np.array(sorted(end_name), dtype=np.int64)
        chosen = chosen[end_stage[chosen] >= 0] if chosen.size else chosen
        if chosen.size == 0:
            return
        pieces: Dict[int, List[np.ndarray]] = {int(j): [] for j in chosen}
        state_rows: Dict[int, np.ndarray] = {}
        first = True
        for si, stage in enumerate(self.stages):
            here = np.array([j for j in chosen if end_stage[j] >= si], dtype=np.int64)      # still in flight at this stage
            if here.size == 0:
                break
            if isinstance(stage, _DeviceStage):
                if first:
                    sel = torch.from_numpy(here).to(self.tdev)
                    rows, off, _ = stage.prop.trajectories(ic, select=sel)
                    skip = 0
                else:
                    st = np.ascontiguousarray(np.stack([state_rows[int(j)] for j in here], axis=1))
                    rows, off, _ = stage.prop.trajectories(torch.from_numpy(st).to(self.tdev))
                    skip = 1                                   # row 0 repeats the row the molecule resumed from
                for k, j in enumerate(here):
                    block = rows[int(off[k]) + skip:int(off[k + 1])]
                    pieces[int(j)].append(np.array(block))
                    state_rows[int(j)] = np.array(rows[int(off[k + 1]) - 1])
            else:
                if first:
                    init = _initial_rows(ic[:, torch.from_numpy(here).to(self.tdev)].cpu().numpy())
                    for k, j in enumerate(here):
                        pieces[int(j)].append(init[k:k + 1])
                for j in here:
                    frag = fragments[si][int(j)]
                    pieces[int(j)].append(frag)
                    if len(frag):
                        state_rows[int(j)] = frag[-1]
                    elif first:
                        state_rows[int(j)] = pieces[int(j)][0][0]
            first = False
        for j in chosen:
            rows = np.concatenate(pieces[int(j)], axis=0)
            rows = rows[np.isfinite(rows).all(axis=1)]          # Trajectory.drop_nans (beamline.py:38)
            name = end_name[int(j)]
            self.molecules.append(Molecule.from_rows(rows, name, name == "Detected"))


def draw_ic(source, seed: int, first: int, n: int, device: int):
    """Philox samples [6, n] on the device (no beamline needed)."""
    torch = eng._torch()
    ic = torch.empty((6, n), dtype=torch.float64, device=torch.device("cuda", device))
    with torch.cuda.device(device):
        nat.check(nat.lib().cmt_philox_draw(C.byref(source), int(seed) & (2**64 - 1), int(first), None, n,
                                            ic.data_ptr(), max(n, 1), eng._stream_ptr(device)))
    return ic


def merge_counts_across_ranks(counts: Dict[str, int]) -> Dict[str, int]:
    """Sum name -> count dictionaries over torch.distributed ranks (names may differ between ranks: a user element is
    free to invent fate names)."""
    torch = eng._torch()
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()) or torch.distributed.get_world_size() == 1:
        return counts
    boxes: List[Optional[dict]] = [None] * torch.distributed.get_world_size()
    torch.distributed.all_gather_object(boxes, counts)
    total: Dict[str, int] = {}
    for box in boxes:
        for name, c in box.items():
            total[name] = total.get(name, 0) + int(c)
    return total
