"""Host-side orchestration of the CUDA propagation path.

PyTorch is plumbing only: device memory (tensors), the current CUDA stream and
`torch.distributed` for the tiny Counter all-reduce.  All arithmetic on the
path happens in libcmt_b200.so (csrc/), reached through the C ABI in
include/cmt.h.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as nat

G = 9.80665  # scipy.constants.g (molecule.py:6)
DEFAULT_CHUNK = 1 << 26          # molecules per launch: larger launches keep the lens integrator's lanes refilled (1e10
                                 # molecules: 0.81 s at 2^24, 0.73 s at 2^26).  The lens-queue workspace of a launch is
                                 # 2 x 64 B per queue entry: 8.6 GB per stream slot if the queue is sized for every
                                 # molecule, 0.13 GB when it is sized from a pilot launch (Propagator.queue_capacity)
PILOT_MOLECULES = 1 << 21        # first launch of a large Philox run: full-size queue, its lens-entry count sizes the rest
SWEEP_PILOT_MOLECULES = 1 << 18  # run_sweep: a throw-away launch of the first point that sizes the queues of all points
ROW_BUDGET_BYTES = 1 << 30       # device bytes per saved-trajectory batch
PINNED_RESULT_BYTES = 2 << 30    # saved-trajectory blocks up to this size are returned in page-locked memory


def _torch():
    import torch

    return torch


# ---------------------------------------------------------------------------
# flattening: Beamline -> element table + lens tables + fate names
# ---------------------------------------------------------------------------
@dataclass
class FlatBeamline:
    elements: List[nat.Element]
    tables: List[Tuple[np.ndarray, np.ndarray]]
    fate_names: List[str]
    max_rows: int
    key: bytes = b""

    @property
    def fate_detected(self) -> int:
        return self.fate_names.index("Detected")

    def fate_id(self, name: str) -> Optional[int]:
        return self.fate_names.index(name) if name in self.fate_names else None

    def save_mask(self, names: Sequence[str]) -> int:
        mask = 0
        for nm in names:
            k = self.fate_id(nm)
            if k is not None:
                mask |= 1 << k
        return mask


def flatten(elements: Sequence) -> FlatBeamline:
    """Flatten beamline elements (already sorted by z0, beamline.py:40-45).

    Fate strings follow the reference: an aperture's `name`
    (apertures.py:113,188,242,264), the literals "Lens entrance"/"Inside lens"
    for a lens (electrostatic_lens.py:63,117) and "Detected" (beamline.py:34-35).
    """
    from .beamline_elements.apertures import CircularAperture, FieldPlates, RectangularAperture
    from .beamline_elements.electrostatic_lens import ElectrostaticLens
    from .beamline_elements.meshes import Honeycomb

    if len(elements) > nat.CMT_MAX_ELEMENTS:
        raise ValueError(f"at most {nat.CMT_MAX_ELEMENTS} beamline elements are supported")
    names: List[str] = []

    def fid(name: str) -> int:
        if name not in names:
            names.append(name)
        return names.index(name)

    out: List[nat.Element] = []
    tables: List[Tuple[np.ndarray, np.ndarray]] = []
    max_rows = 1
    for e in elements:
        t = nat.Element()
        t.z0, t.z1 = float(e.z0), float(e.z1)
        if isinstance(e, CircularAperture):
            t.type, t.fate, t.R = nat.CIRCULAR, fid(e.name), float(e.d) / 2
            max_rows += 2
        elif isinstance(e, RectangularAperture):
            t.type, t.fate = nat.RECTANGULAR, fid(e.name)
            t.x1, t.x2, t.y1, t.y2 = float(e.x1), float(e.x2), float(e.y1), float(e.y2)
            max_rows += 2
        elif isinstance(e, FieldPlates):
            t.type, t.fate = nat.FIELDPLATES, fid(e.name)
            t.x1, t.x2 = float(e.x1), float(e.x2)
            max_rows += 2
        elif isinstance(e, ElectrostaticLens):
            t.type = nat.LENS
            t.fate, t.fate2 = fid("Lens entrance"), fid("Inside lens")
            t.R, t.dz = float(e.d) / 2, float(e.dz)
            t.n_steps = int(np.rint(e.L / e.dz))
            r, a = e.acceleration_table()
            t.table = len(tables)
            tables.append((np.ascontiguousarray(r, dtype=np.float64), np.ascontiguousarray(a, dtype=np.float64)))
            max_rows += 2 + t.n_steps
        elif isinstance(e, Honeycomb):
            t.type, t.fate = nat.HONEYCOMB, fid(e.name)
            t.R, t.dz = e.polygon_radius, e.pitch
            t.x1, t.y1 = e.grid_origin()
            t.n_steps, t.reserved = int(e.nx), int(e.ny)
            max_rows += 2
        else:
            raise TypeError(
                f"beamline element {type(e).__name__!r} has no CUDA implementation "
                "(supported: CircularAperture, RectangularAperture, FieldPlates, ElectrostaticLens, Honeycomb)"
            )
        out.append(t)
    fid("Detected")
    if len(names) > nat.CMT_MAX_FATES:
        raise ValueError(f"at most {nat.CMT_MAX_FATES} distinct fates are supported")
    if len(tables) > nat.CMT_MAX_TABLES:
        raise ValueError(f"at most {nat.CMT_MAX_TABLES} lenses are supported")
    key = b"".join(bytes(t) for t in out) + b"|".join(r.tobytes() + a.tobytes() for r, a in tables)
    key += "|".join(names).encode()
    return FlatBeamline(out, tables, names, max_rows, key)


# ---------------------------------------------------------------------------
# device handle
# ---------------------------------------------------------------------------
class DeviceBeamline:
    """Owns a cmt_beamline_t on one GPU."""

    def __init__(self, flat: FlatBeamline, device: int, math: str = "exact"):
        lib = nat.lib()
        self.flat = flat
        self.device = int(device)
        if math not in nat.MATH_MODES:
            raise ValueError(f"math must be one of {sorted(nat.MATH_MODES)}, got {math!r}")
        self.math = math
        arr = (nat.Element * max(len(flat.elements), 1))(*flat.elements)
        tabs = (nat.Table * max(len(flat.tables), 1))()
        for k, (r, a) in enumerate(flat.tables):
            tabs[k].r = r.ctypes.data_as(C.POINTER(C.c_double))
            tabs[k].a = a.ctypes.data_as(C.POINTER(C.c_double))
            tabs[k].n = len(r)
        handle = C.c_void_p()
        nat.check(lib.cmt_beamline_create(arr, len(flat.elements), tabs, len(flat.tables), len(flat.fate_names),
                                          flat.fate_detected, G, self.device, C.byref(handle)))
        self.handle = handle
        nat.check(lib.cmt_beamline_set_math(handle, nat.MATH_MODES[math]))
        self.max_rows = lib.cmt_beamline_max_rows(handle)

    def workspace_bytes(self, n: int) -> int:
        return int(nat.lib().cmt_workspace_bytes(self.handle, int(n)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                nat.lib().cmt_beamline_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_handles: Dict[Tuple[bytes, int, str], DeviceBeamline] = {}


def device_beamline(flat: FlatBeamline, device: int, math: str = "exact") -> DeviceBeamline:
    k = (flat.key, int(device), math)
    h = _handles.get(k)
    if h is None:
        if len(_handles) > 256:
            _handles.clear()
        h = _handles[k] = DeviceBeamline(flat, device, math)
    return h


def resolve_device(device=None) -> int:
    torch = _torch()
    if not torch.cuda.is_available():
        raise nat.NativeError("no CUDA device is available and there is no CPU fallback for the propagation path")
    if device is None:
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    d = torch.device(device)
    return d.index if d.index is not None else torch.cuda.current_device()


def _stream_ptr(device: int) -> int:
    return int(_torch().cuda.current_stream(device).cuda_stream)


# ---------------------------------------------------------------------------
# source description
# ---------------------------------------------------------------------------
def make_source(vdist, xdist) -> Optional[nat.Source]:
    """Distribution objects -> device source record, or None when the pair has
    no on-device generator (then draw() is called on the host and replayed)."""
    from .distributions import (CeNTREXPositionDistribution, CeNTREXVelocityDistribution,
                                GaussianPositionDistribution)

    if type(vdist) is not CeNTREXVelocityDistribution:
        return None
    s = nat.Source()
    s.vmean[:] = [vdist.vx, vdist.vy, vdist.vz]
    s.vsigma[:] = [vdist.sigmax, vdist.sigmay, vdist.sigmaz]
    if type(xdist) is CeNTREXPositionDistribution:
        s.pos_kind, s.p0, s.p1 = nat.POS_DISC, xdist.d / 2, 0.0
    elif type(xdist) is GaussianPositionDistribution:
        s.pos_kind, s.p0, s.p1 = nat.POS_GAUSS, xdist.sigmax, xdist.sigmay
    else:
        return None
    s.z = xdist.z
    return s


# ---------------------------------------------------------------------------
# one propagation call over device-resident data
# ---------------------------------------------------------------------------
@dataclass
class PropagateResult:
    counters: "object"                     # torch int64 [n_fates] (device)
    work: "object"                         # torch int64 [4] (device)
    fate: Optional["object"] = None        # torch uint8 [n]
    final: Optional["object"] = None       # torch float64 [10, n]
    saved_index: Optional["object"] = None  # torch int64 [n_saved], sorted
    fate_names: List[str] = field(default_factory=list)


class GraphedStep:
    """A captured propagation step: `replay()` enqueues it on its stream."""

    def __init__(self, graph, stream, fate, keep):
        self.graph, self.stream, self.fate, self._keep = graph, stream, fate, keep

    def replay(self):
        torch = _torch()
        with torch.cuda.stream(self.stream):
            self.graph.replay()


_SLOT_STREAMS: dict = {}      # device index -> the side streams every Propagator of that device launches on


class Propagator:
    """Reusable launch context: beamline handle + scratch buffers on one device.

    Launches normally go to the current CUDA stream.  With `slot=k` a launch goes to one of
    `n_slots` private streams (each with its own workspace), so consecutive chunks overlap:
    the lens integrator of chunk i, which cannot fill the chip on its own at 1e7 molecules,
    runs beside the walk kernel of chunk i+1.  `join()` makes the current stream wait for them.
    """

    def __init__(self, elements_or_flat, device=None, n_slots: int = 3, math: str = "exact"):
        self.flat = elements_or_flat if isinstance(elements_or_flat, FlatBeamline) else flatten(elements_or_flat)
        self.device = resolve_device(device)
        self.math = math
        self.dev = device_beamline(self.flat, self.device, math)
        torch = _torch()
        self.tdev = torch.device("cuda", self.device)
        self.counters = torch.zeros(len(self.flat.fate_names), dtype=torch.int64, device=self.tdev)
        self.work = torch.zeros(8, dtype=torch.int64, device=self.tdev)
        self.n_slots = int(n_slots)
        self._ws = [None] * (self.n_slots + 1)           # last entry: launches on the current stream
        self._saved_count = [torch.zeros(1, dtype=torch.int64, device=self.tdev) for _ in range(self.n_slots + 1)]
        self._streams = None
        self.entry_fraction = None                       # lens entries per molecule seen by a pilot launch (None: unknown)

    # -- lens-queue sizing -------------------------------------------------------
    def queue_capacity(self, n: int) -> int:
        """Lens-queue entries for a launch of n molecules: all n while nothing is known about the source; after
        `learn_entry_fraction()` the expected number of molecules that reach the lens plus 25 % and six standard
        deviations, at least n / 64.  A launch that overflows its queue drops molecules and says so in work[6];
        `queue_overflow()` reads it."""
        if self.entry_fraction is None:
            return int(n)
        mean = self.entry_fraction * n
        return int(min(n, max(n // 64, 1.25 * mean + 6.0 * mean ** 0.5 + 4096)))

    def learn_entry_fraction(self, n_launched: int) -> float:
        """Read the work counters (one synchronisation) after `n_launched` molecules and size later queues from them."""
        w = self.work.cpu().numpy()
        self.entry_fraction = float(w[3]) / max(int(n_launched), 1)
        return self.entry_fraction

    def queue_overflow(self) -> int:
        return int(self.work[6].item())

    def reset(self):
        self.join()
        self.counters.zero_()
        self.work.zero_()

    def rebind(self, flat: FlatBeamline):
        """Point this launch context at another beamline with the same fates (a sweep point: same elements, another
        lens table) and give it fresh Counter / work tensors; streams and workspaces are kept, and launches already
        queued for the previous beamline keep their own handle and counters."""
        if flat.fate_names != self.flat.fate_names:
            raise ValueError("rebind needs a beamline with the same fates")
        torch = _torch()
        self.flat = flat
        self.dev = device_beamline(flat, self.device, self.math)
        self.counters = torch.zeros(len(flat.fate_names), dtype=torch.int64, device=self.tdev)
        self.work = torch.zeros(8, dtype=torch.int64, device=self.tdev)

    def release(self):
        """Give the queue workspaces back to the allocator."""
        self.join()
        _torch().cuda.current_stream(self.device).synchronize()
        self._ws = [None] * (self.n_slots + 1)

    def fit_chunk(self, chunk: int, fraction: float = 0.5) -> int:
        """Largest launch size <= `chunk` whose workspaces (one per stream slot plus one, 128 B per molecule each, and
        the saved-index buffer) fit into `fraction` of the device memory that is free right now."""
        free, _total = _torch().cuda.mem_get_info(self.device)
        held = sum(w.numel() for w in self._ws if w is not None)
        queue_share = self.queue_capacity(1 << 26) / float(1 << 26)
        per_molecule = (self.n_slots + 1) * 2 * 8 * 8 * queue_share + 16
        fit = int(fraction * (free + held)) // per_molecule
        return max(1 << 20, min(int(chunk), fit))

    def dev_sm_count(self) -> int:
        return int(_torch().cuda.get_device_properties(self.device).multi_processor_count)

    # -- streams ---------------------------------------------------------------
    def _slot_stream(self, slot):
        # The slot streams belong to the device, not to the Propagator: run_simulation makes a Propagator per call,
        # and the caching allocator keeps freed blocks per stream -- with fresh streams every call the 128 B x n
        # workspaces of a 1e7-molecule run were allocated from the driver again each time (0.5 ms per
        # torch.empty, 1 ms of a 2 ms call; profiles/prof_api.py).
        torch = _torch()
        if self._streams is None:
            pool = _SLOT_STREAMS.setdefault(self.device, [])
            with torch.cuda.device(self.device):
                while len(pool) < self.n_slots:
                    pool.append(torch.cuda.Stream(device=self.tdev))
            self._streams = pool[:self.n_slots]
        return self._streams[slot % self.n_slots]

    def join(self):
        """Make the current stream wait for every slot stream."""
        if self._streams is not None:
            cur = _torch().cuda.current_stream(self.device)
            for st in self._streams:
                cur.wait_stream(st)

    class _Launch:
        def __init__(self, prop, slot):
            self.prop, self.slot, self.ctx = prop, slot, None

        def __enter__(self):
            torch = _torch()
            p = self.prop
            self.caller = torch.cuda.current_stream(p.device)
            if self.slot is None:
                self.index = p.n_slots
                self.ctx = torch.cuda.device(p.device)
            else:
                st = p._slot_stream(self.slot)
                st.wait_stream(self.caller)                           # inputs produced on the caller's stream
                self.index = self.slot % p.n_slots
                self.ctx = torch.cuda.stream(st)
            self.ctx.__enter__()
            return self

        def __exit__(self, *exc):
            return self.ctx.__exit__(*exc)

        def hand_over(self, *tensors):
            """Outputs allocated on a slot stream are consumed on the caller's stream (after join()): tell the
            caching allocator, so that their memory is not reused while the caller's work is still queued."""
            if self.slot is not None:
                for t in tensors:
                    if t is not None:
                        t.record_stream(self.caller)

    def _workspace(self, n: int, index: int):
        torch = _torch()
        need = self.dev.workspace_bytes(self.queue_capacity(n))
        if self._ws[index] is None or self._ws[index].numel() < need:
            self._ws[index] = torch.empty(need, dtype=torch.uint8, device=self.tdev)
        return self._ws[index]

    def _outputs(self, n, want_fate, want_final, save_mask, saved_buf, index):
        torch = _torch()
        O = nat.Outputs()
        fate = final = None
        if want_fate:
            fate = torch.empty(n, dtype=torch.uint8, device=self.tdev)
            O.fate = fate.data_ptr()
        if want_final:
            final = torch.empty((10, n), dtype=torch.float64, device=self.tdev)
            O.final_state, O.final_ld = final.data_ptr(), n
        O.counters = self.counters.data_ptr()
        O.work = self.work.data_ptr()
        cap = self.queue_capacity(n)
        O.queue_capacity = cap if cap < n else 0
        if save_mask:
            self._saved_count[index].zero_()
            O.saved_index = saved_buf.data_ptr()
            O.saved_count = self._saved_count[index].data_ptr()
            O.saved_capacity = saved_buf.numel()
            O.save_mask = save_mask
        return O, fate, final

    def _finish_saved(self, save_mask, saved_buf, index):
        if not save_mask:
            return None
        torch = _torch()
        k = int(self._saved_count[index].item())
        if k > saved_buf.numel():
            raise RuntimeError("saved-index buffer overflow")  # cannot happen: capacity == n
        return torch.sort(saved_buf[:k]).values

    def propagate_ic(self, ic, first_index=0, want_fate=True, want_final=False, save_mask=0,
                     slot=None) -> PropagateResult:
        """ic: torch float64 [6, n] on this device (SoA x,y,z,vx,vy,vz)."""
        torch = _torch()
        assert ic.dtype == torch.float64 and ic.dim() == 2 and ic.shape[0] == 6 and ic.is_cuda
        assert ic.shape[1] == 0 or ic.stride(1) == 1
        n = ic.shape[1]
        with self._Launch(self, slot) as L:
            ws = self._workspace(n, L.index)
            saved_buf = torch.empty(n, dtype=torch.int64, device=self.tdev) if save_mask else None
            O, fate, final = self._outputs(n, want_fate, want_final, save_mask, saved_buf, L.index)
            nat.check(nat.lib().cmt_propagate_ic(self.dev.handle, n, int(first_index), ic.data_ptr(), ic.stride(0),
                                                 C.byref(O), ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
            saved = self._finish_saved(save_mask, saved_buf, L.index)
            L.hand_over(fate, final, saved)
        return PropagateResult(self.counters, self.work, fate, final, saved, self.flat.fate_names)

    def propagate_philox(self, source: nat.Source, seed: int, first_index: int, n: int, want_fate=False,
                         want_final=False, save_mask=0, slot=None) -> PropagateResult:
        torch = _torch()
        with self._Launch(self, slot) as L:
            ws = self._workspace(n, L.index)
            saved_buf = torch.empty(n, dtype=torch.int64, device=self.tdev) if save_mask else None
            O, fate, final = self._outputs(n, want_fate, want_final, save_mask, saved_buf, L.index)
            nat.check(nat.lib().cmt_propagate_philox(self.dev.handle, C.byref(source), int(seed) & (2**64 - 1),
                                                     int(first_index), int(n), C.byref(O), ws.data_ptr(),
                                                     ws.numel(), _stream_ptr(self.device)))
            saved = self._finish_saved(save_mask, saved_buf, L.index)
            L.hand_over(fate, final, saved)
        return PropagateResult(self.counters, self.work, fate, final, saved, self.flat.fate_names)

    def capture_ic(self, ic, first_index=0, want_fate=True, slot=0) -> "GraphedStep":
        """Capture one propagate_ic call (header memset + walk + lens segment + tail kernels) into a CUDA graph on
        the slot's stream.  Replaying it costs one graph launch on the host instead of three
        launches plus the Python call path, which matters when a step lasts well under a
        millisecond and several processes share the host cores."""
        torch = _torch()
        assert ic.dtype == torch.float64 and ic.dim() == 2 and ic.shape[0] == 6 and ic.is_cuda and ic.stride(1) == 1
        n = ic.shape[1]
        index = slot % self.n_slots
        st = self._slot_stream(slot)
        ws = self._workspace(n, index)
        O, fate, _ = self._outputs(n, want_fate, False, 0, None, index)
        st.wait_stream(torch.cuda.current_stream(self.device))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.device(self.device), torch.cuda.graph(graph, stream=st):
            nat.check(nat.lib().cmt_propagate_ic(self.dev.handle, n, int(first_index), ic.data_ptr(), ic.stride(0),
                                                 C.byref(O), ws.data_ptr(), ws.numel(), _stream_ptr(self.device)))
        # the graph holds raw pointers: keep inputs, workspace, outputs and the beamline handle alive with it
        return GraphedStep(graph, st, fate, (ic, ws, O, self.dev, self.counters, self.work))

    def draw(self, source: nat.Source, seed: int, first_index: int = 0, n: int = 0, index=None):
        """Materialise source samples as a device tensor [6, n]."""
        torch = _torch()
        if index is not None:
            n = index.numel()
        ic = torch.empty((6, n), dtype=torch.float64, device=self.tdev)
        with torch.cuda.device(self.device):
            nat.check(nat.lib().cmt_philox_draw(C.byref(source), int(seed) & (2**64 - 1), int(first_index),
                                                index.data_ptr() if index is not None else None, n,
                                                ic.data_ptr(), max(n, 1), _stream_ptr(self.device)))
        return ic

    def _traj_call(self, m, state, n_comp, sel, select_base, rows_ptr, row_offset_ptr, n_rows, fate):
        torch = _torch()
        with torch.cuda.device(self.device):
            nat.check(nat.lib().cmt_trajectories(
                self.dev.handle, m, state.data_ptr(), n_comp, state.stride(0), sel, int(select_base), rows_ptr,
                self.dev.max_rows, row_offset_ptr, n_rows.data_ptr(), fate.data_ptr(), _stream_ptr(self.device)))

    def plane_crossings(self, state, z_planes, select=None, select_base=0):
        """State of the molecules in `state` ([6|10, m] device tensor, optionally gathered through `select`)
        where they cross the planes `z_planes`: the device form of post_processing.find_radial_pos_dist /
        find_vel_dist, with no trajectory stored anywhere.

        Returns device tensors (out [n_planes, 5, k] = x, y, vx, vy, vz; valid [n_planes, k] bool;
        fate [k] uint8), planes in the caller's order."""
        torch = _torch()
        zs = np.atleast_1d(np.asarray(z_planes, dtype=np.float64))
        if zs.ndim != 1 or zs.size == 0:
            raise ValueError("z_planes must be a non-empty 1-d sequence")
        if np.isnan(zs).any():
            raise ValueError("z_planes contains NaN")
        n_comp = state.shape[0]
        k_total = select.numel() if select is not None else state.shape[1]
        assert state.dtype == torch.float64 and (state.stride(1) == 1 or state.shape[1] == 0)
        out = torch.zeros((zs.size, 5, k_total), dtype=torch.float64, device=self.tdev)
        valid = torch.zeros((zs.size, k_total), dtype=torch.uint8, device=self.tdev)
        fate = torch.empty(k_total, dtype=torch.uint8, device=self.tdev)
        if k_total == 0:
            return out, valid.bool(), fate
        order = np.argsort(zs, kind="stable")
        sel_ptr = select.data_ptr() if select is not None else None
        with torch.cuda.device(self.device):
            for lo in range(0, zs.size, nat.CMT_MAX_PLANES):
                idx = order[lo:lo + nat.CMT_MAX_PLANES]
                group = np.ascontiguousarray(zs[idx])
                # sorted planes are written to a contiguous scratch block, then scattered to the caller's order
                o = torch.zeros((idx.size, 5, k_total), dtype=torch.float64, device=self.tdev)
                v = torch.zeros((idx.size, k_total), dtype=torch.uint8, device=self.tdev)
                nat.check(nat.lib().cmt_plane_crossings(
                    self.dev.handle, k_total, state.data_ptr(), n_comp, state.stride(0), sel_ptr, int(select_base),
                    group.ctypes.data_as(C.POINTER(C.c_double)), int(idx.size), o.data_ptr(), k_total,
                    v.data_ptr(), fate.data_ptr(), _stream_ptr(self.device)))
                where = torch.from_numpy(idx.astype(np.int64)).to(self.tdev)
                out[where] = o
                valid[where] = v
        return out, valid.bool(), fate

    def trajectories(self, state, select=None, select_base=0, defer_sync=False):
        """Full trajectories of the molecules in `state` ([6|10, m] device tensor), optionally gathered
        through `select` (global indices, device int64).

        Two passes per batch: a counting call gives every molecule's row count, an exclusive scan turns
        the counts into row offsets, and the second call writes the rows compactly, so memory is
        sum(n_rows) x 80 B whatever the fates are (a molecule stopped by the first aperture has 2 rows,
        a detected one 613).  Rows come back through two reusable pinned staging buffers.

        Returns (rows [total_rows, 10], offsets [k + 1] int64, fate [k]) as host arrays; molecule j owns
        rows[offsets[j]:offsets[j + 1]].  With `defer_sync` a fourth item follows, a callable that waits for the rows
        to have arrived: offsets and fates are final after the first pass, so the caller can build its per-molecule
        objects (views of `rows`, 2 ms for 2385 molecules) while the device-to-host copy (2 ms for their 117 MB) is
        still in flight, and call it before reading a row.
        """
        torch = _torch()
        n_comp = state.shape[0]
        k_total = select.numel() if select is not None else state.shape[1]
        row_budget = max(self.dev.max_rows, ROW_BUDGET_BYTES // (nat.CMT_ROW_DOUBLES * 8))

        # pass 1: counts and fates for everything
        n_rows = torch.empty(k_total, dtype=torch.int32, device=self.tdev)
        fate = torch.empty(k_total, dtype=torch.uint8, device=self.tdev)
        if k_total:
            sel_ptr = select.data_ptr() if select is not None else None
            self._traj_call(k_total, state, n_comp, sel_ptr, select_base, None, None, n_rows, fate)
        offsets = torch.zeros(k_total + 1, dtype=torch.int64, device=self.tdev)
        torch.cumsum(n_rows, 0, out=offsets[1:])
        off_np = offsets.cpu().numpy()
        fate_np = fate.cpu().numpy()        # final after the counting pass (the second pass writes the same bytes)
        total_rows = int(off_np[-1])
        # Result block.  Up to PINNED_RESULT_BYTES it is page-locked memory from a small pool (_pinned_rows): the
        # device copies straight into the array the caller receives (no staging, no host-side copy, no first-touch
        # page faults), and the block is free for the next run once the last trajectory viewing it has been dropped,
        # so a loop of runs re-uses the same pages.  Larger results, results that find the pool busy, or a failed
        # page-lock go through two reusable pinned staging buffers into ordinary memory.
        host = rows_out = None
        if 0 < total_rows * nat.CMT_ROW_DOUBLES * 8 <= PINNED_RESULT_BYTES:
            host, rows_out = _pinned_rows(total_rows)
        if host is None:
            rows_out = np.empty((total_rows, nat.CMT_ROW_DOUBLES), dtype=np.float64)

        # pass 2: rows, in batches bounded by the device row budget
        stage = _staging(self.device) if host is None else None
        stage_rows = stage[0].shape[0] if stage is not None else 0
        pending = None                     # (event, staging index, destination slice)

        def drain():
            nonlocal pending
            if pending is not None:
                ev, b, lo_r, hi_r = pending
                ev.synchronize()
                rows_out[lo_r:hi_r] = stage[b][: hi_r - lo_r].numpy()
                pending = None

        lo, piece = 0, 0
        while lo < k_total:
            hi = int(np.searchsorted(off_np, off_np[lo] + row_budget, side="right")) - 1
            hi = min(max(hi, lo + 1), k_total)
            base, n_batch_rows = int(off_np[lo]), int(off_np[hi] - off_np[lo])
            rows = torch.empty((max(n_batch_rows, 1), nat.CMT_ROW_DOUBLES), dtype=torch.float64, device=self.tdev)
            rel = offsets[lo:hi] - base
            if select is not None:
                sel_ptr, st = select[lo:hi].data_ptr(), state
            else:
                sel_ptr, st = None, state[:, lo:hi]
            self._traj_call(hi - lo, st, n_comp, sel_ptr, select_base, rows.data_ptr(), rel.data_ptr(),
                            n_rows[lo:hi], fate[lo:hi])
            if host is not None:
                if n_batch_rows:
                    host[base:base + n_batch_rows].copy_(rows[:n_batch_rows], non_blocking=True)
                lo = hi
                continue
            # each piece comes back through a pinned staging buffer, the D2H of one piece overlapping
            # the host copy-out of the previous one
            for r0 in range(0, n_batch_rows, stage_rows):
                r1 = min(n_batch_rows, r0 + stage_rows)
                b = piece & 1
                if pending is not None and pending[1] == b:
                    drain()
                stage[b][: r1 - r0].copy_(rows[r0:r1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                prev, pending = pending, (ev, b, base + r0, base + r1)
                if prev is not None:           # copy the previous piece out while this one is in flight
                    pev, pb, plo, phi = prev
                    pev.synchronize()
                    rows_out[plo:phi] = stage[pb][: phi - plo].numpy()
                piece += 1
            drain()                            # the device buffer is released before the next batch
            lo = hi
        if host is None:
            wait = lambda: None                                 # the staged path has copied everything out already
        else:
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            wait = done.synchronize
        if defer_sync:
            return rows_out, off_np, fate_np, wait
        wait()
        return rows_out, off_np, fate_np


# Page-locked result blocks.  Locking pages costs ~0.4 ms per MB (45 ms for the 117 MB of a 1e7-molecule run of the lens
# beamline), far more than the copy it speeds up, so it only pays for memory that is used again: a loop of runs whose
# previous result has been dropped.  The pool keeps at most PINNED_POOL_BLOCKS blocks; a block is free again when the
# array handed out over it -- the base of every trajectory view of that result -- has been collected.  A result that
# finds no free block while the pool is full (results that are all kept: the 16 launches of configs[3], a loop over
# 40 sweep points) goes through the pinned staging buffers into ordinary memory instead, which is faster than locking
# fresh pages for it.
PINNED_POOL_BLOCKS = 3
PINNED_BLOCK_GRAIN = 32 << 20
_PINNED_POOL: list = []       # [block (uint8 tensor, page-locked), weakref to the array last handed out over it or None]
_PINNED_LOCK = threading.Lock()


def _pinned_rows(total_rows: int):
    """(tensor [total_rows, 10] float64 over page-locked memory, the same memory as a NumPy array) or (None, None)."""
    import weakref

    torch = _torch()
    nbytes = total_rows * nat.CMT_ROW_DOUBLES * 8
    with _PINNED_LOCK:          # two threads running simulations must not be handed the same free block
        return _pinned_rows_locked(torch, weakref, total_rows, nbytes)


def _pinned_rows_locked(torch, weakref, total_rows: int, nbytes: int):
    free = [e for e in _PINNED_POOL if e[1] is None or e[1]() is None]
    fit = [e for e in free if e[0].numel() >= nbytes]
    entry = min(fit, key=lambda e: e[0].numel()) if fit else None
    if entry is None and (len(_PINNED_POOL) < PINNED_POOL_BLOCKS or free):
        if len(_PINNED_POOL) >= PINNED_POOL_BLOCKS:
            victim = max(free, key=lambda e: e[0].numel())                  # too small: make room for a larger one
            _PINNED_POOL[:] = [e for e in _PINNED_POOL if e is not victim]
        capacity = -(-nbytes // PINNED_BLOCK_GRAIN) * PINNED_BLOCK_GRAIN
        try:
            entry = [torch.empty(capacity, dtype=torch.uint8, pin_memory=True), None]
        except RuntimeError:
            return None, None
        _PINNED_POOL.append(entry)
    if entry is None:
        return None, None
    host = entry[0][:nbytes].view(torch.float64).view(total_rows, nat.CMT_ROW_DOUBLES)
    rows = host.numpy()
    entry[1] = weakref.ref(rows)
    return host, rows


_STAGING: Dict[int, list] = {}
STAGING_BYTES = 32 << 20


def _staging(device: int):
    """Two pinned staging buffers per device, allocated once per process."""
    if device not in _STAGING:
        torch = _torch()
        n = STAGING_BYTES // (nat.CMT_ROW_DOUBLES * 8)
        _STAGING[device] = [torch.empty((n, nat.CMT_ROW_DOUBLES), dtype=torch.float64, pin_memory=True) for _ in range(2)]
    return _STAGING[device]


# ---------------------------------------------------------------------------
# distributed helpers (one process per GPU; NCCL on GPUs, gloo in CPU tests)
# ---------------------------------------------------------------------------
def dist_info() -> Tuple[int, int]:
    torch = _torch()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of the global index range [0,total) owned by `rank`."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owned_draws(vdist, xdist, N: int, N_loops: int, rank: int, world: int):
    """The (velocities, positions) chunks of the reference's draw loop (trajectory_simulator.py:57-58) that belong
    to `rank`: a contiguous block of the N_loops chunks.  Every rank draws from chunk 0 on and discards the
    chunks of lower ranks, so ranks whose generators are seeded identically end up with disjoint pieces of one
    sample (see run_simulation)."""
    if N == 0:
        return
    lo_loop, hi_loop = shard_range(N_loops, rank, world)
    for loop in range(hi_loop):
        vs = np.asarray(vdist.draw(N), dtype=np.float64)   # velocities first, as the reference
        xs = np.asarray(xdist.draw(N), dtype=np.float64)
        if vs.shape != (3, N) or xs.shape != (3, N):
            raise ValueError(f"Distribution.draw({N}) must return shape (3, {N})")
        if loop >= lo_loop:
            yield vs, xs


def allreduce_counts(t):
    """Sum an int64 tensor over ranks (the Counter merge, trajectory_simulator.py:147-158)."""
    torch = _torch()
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
    return t


def gather_counts(n: int, device=None) -> List[int]:
    """all_gather of one int64 per rank (the saved-trajectory counts): every rank learns how many
    molecules each rank saved, hence the global number of its own first saved molecule."""
    torch = _torch()
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()) or torch.distributed.get_world_size() == 1:
        return [int(n)]
    dist = torch.distributed
    dev = torch.device("cuda", device) if (dist.get_backend() == "nccl" and device is not None) else torch.device("cpu")
    mine = torch.tensor([int(n)], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [int(t.item()) for t in out]


def broadcast_object(obj, src: int = 0):
    """Small Python object from rank `src` to every rank (a no-op without torch.distributed)."""
    torch = _torch()
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()) or torch.distributed.get_world_size() == 1:
        return obj
    box = [obj]
    torch.distributed.broadcast_object_list(box, src=src)
    return box[0]


def barrier():
    torch = _torch()
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        torch.distributed.barrier()
