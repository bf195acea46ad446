"""Molecule / Trajectory containers (API of the reference's molecule.py:8-186).

On the GPU path a molecule is ten FP64 registers; these classes are the host
view of a finished (or hand-built) trajectory: `x`, `v`, `a` are (n,3) float64
arrays and `t` is (n,), which is what post-processing and the HDF layout
expect.  Trajectories that come back from the GPU are zero-copy views of one
[n,10] row block (columns x,y,z,vx,vy,vz,ax,ay,az,t).
"""
from __future__ import annotations

from bisect import bisect_right
from collections.abc import MutableSequence
from dataclasses import dataclass

import numpy as np

g = 9.80665  # scipy.constants.g
_DEFAULT_A = np.array((0, -g, 0))


class Trajectory:
    """Row store: one row per recorded point, `n` rows in use."""

    __slots__ = ("x", "v", "a", "t", "n", "_rows")

    def __init__(self, beamline=None, n_rows: int = None):
        if n_rows is None:
            # the reference sizes the arrays as 10 spare rows + every element's N_steps()
            n_rows = 10 + sum(element.N_steps() for element in beamline.elements)
        self._allocate(n_rows)
        self.n = 0

    def _allocate(self, n_rows: int) -> None:
        nan = np.nan
        self.x, self.v, self.a = (np.full((n_rows, 3), nan) for _ in range(3))
        self.t = np.full(n_rows, nan)

    @classmethod
    def from_rows(cls, rows: np.ndarray) -> "Trajectory":
        """Zero-copy view of a row block [n,10]."""
        tr = object.__new__(cls)
        tr._rows = rows          # x, v, a, t are sliced out of it on first access (__getattr__)
        tr.n = rows.shape[0]
        return tr

    def __getattr__(self, name):
        # Only reached while a slot is unset: a trajectory wrapped by from_rows materialises its four
        # views the first time one of them is asked for (a run that returns thousands of molecules
        # otherwise spends more time slicing than propagating); afterwards they are plain attributes.
        if name in ("x", "v", "a", "t"):
            try:
                rows = object.__getattribute__(self, "_rows")
            except AttributeError:
                raise AttributeError(name) from None
            self.x, self.v, self.a, self.t = rows[:, 0:3], rows[:, 3:6], rows[:, 6:9], rows[:, 9]
            return object.__getattribute__(self, name)
        raise AttributeError(name)

    def as_rows(self) -> np.ndarray:
        """The used rows as one [n,10] block."""
        k = self.n
        return np.concatenate([self.x[:k], self.v[:k], self.a[:k], self.t[:k, None]], axis=1)

    def _grow(self, extra: int) -> None:
        pad3, pad1 = np.full((extra, 3), np.nan), np.full(extra, np.nan)
        self.x, self.v, self.a = (np.concatenate((arr, pad3)) for arr in (self.x, self.v, self.a))
        self.t = np.concatenate((self.t, pad1))

    def update(self, x, v, a, t) -> None:
        """Append one row."""
        if self.n == self.t.shape[0]:
            self._grow(max(16, self.n))
        k = self.n
        self.x[k], self.v[k], self.a[k], self.t[k] = x, v, a, t
        self.n = k + 1

    def extend_rows(self, rows: np.ndarray) -> None:
        """Append a block of GPU-produced rows [k,10]."""
        k = rows.shape[0]
        short = self.n + k - self.t.shape[0]
        if short > 0:
            self._grow(short)
        where = slice(self.n, self.n + k)
        self.x[where], self.v[where], self.a[where], self.t[where] = rows[:, 0:3], rows[:, 3:6], rows[:, 6:9], rows[:, 9]
        self.n += k

    def add_steps(self, beamline) -> None:
        self._grow(sum(element.N_steps() for element in beamline.elements))

    def drop_nans(self) -> None:
        """Trim the unused (NaN) tail of every array."""
        keep3 = lambda arr: arr[np.isfinite(arr).all(axis=1)]  # noqa: E731
        self.x, self.v, self.a = keep3(self.x), keep3(self.v), keep3(self.a)
        self.t = self.t[np.isfinite(self.t)]

    def save_to_hdf(self, file, run_name: str, group_name: str) -> None:
        """Datasets x, v, a, t under <run>/<group> (reference layout)."""
        self.drop_nans()
        group = file.create_group(run_name + "/" + group_name)
        for key in ("x", "v", "a", "t"):
            group.create_dataset(key, data=getattr(self, key))


@dataclass
class Molecule:
    alive: bool = True

    # -- construction ---------------------------------------------------------
    def init_trajectory(self, beamline, x0=np.array((0, 0, 0)), v0=np.array((0, 0, 200)), a0=_DEFAULT_A, t0=0):
        self.trajectory = Trajectory(beamline)
        self.trajectory.update(x0, v0, a0, t0)

    @classmethod
    def from_rows(cls, rows: np.ndarray, aperture_hit: str, alive: bool) -> "Molecule":
        """Wrap GPU-produced rows [n,10]."""
        tr = object.__new__(Trajectory)
        tr._rows = rows
        tr.n = rows.shape[0]
        mol = object.__new__(cls)      # same state as cls(alive=alive) followed by the two assignments
        mol.__dict__ = {"alive": alive, "trajectory": tr, "aperture_hit": aperture_hit}
        return mol

    # -- state of the last recorded point --------------------------------------
    def _last(self, name: str):
        tr = self.trajectory
        return getattr(tr, name)[tr.n - 1]

    def a(self):
        return self._last("a")

    def t(self):
        return self._last("t")

    def x(self, delta_t: float = None):
        """Position now, or after a further ballistic flight of delta_t (falsy delta_t: now)."""
        here = self._last("x")
        if not delta_t:
            return here
        return here + self._last("v") * delta_t + self._last("a") * delta_t ** 2 / 2

    def v(self, delta_t: float = None):
        """Velocity now, or after a further ballistic flight of delta_t."""
        now = self._last("v")
        if not delta_t:
            return now
        return now + self._last("a") * delta_t

    def update_trajectory(self, delta_t, a=_DEFAULT_A):
        """Record the point reached after delta_t; the new row stores `a` (default gravity)."""
        self.trajectory.update(self.x(delta_t), self.v(delta_t), a, self.t() + delta_t)

    # -- fate -----------------------------------------------------------------
    def set_aperture_hit(self, aperture_name):
        self.aperture_hit = aperture_name

    def set_dead(self):
        self.alive = False

    # -- output ---------------------------------------------------------------
    def plot_trajectory(self, axes):
        colour = {"Detected": "g", "Field plates": "r"}.get(self.aperture_hit, "k")
        xs = self.trajectory.x
        axes[0].plot(xs[:, 2], xs[:, 0], c=colour)
        axes[1].plot(xs[:, 2], xs[:, 1], c=colour)

    def save_to_hdf(self, file, run_name: str, group_name: str):
        self.trajectory.save_to_hdf(file, run_name, group_name)
        attrs = file[run_name + "/" + group_name].attrs
        attrs["aperture_hit"] = self.aperture_hit
        attrs["alive"] = self.alive


class SavedMolecules(MutableSequence):
    """The saved molecules of a run, in the order the reference returns them (trajectory_simulator.py:86-91), as a
    list-like sequence whose `Molecule` objects are made the first time they are asked for.

    A GPU run hands back one block of rows per launch plus, per molecule, its slice and its fate; wrapping every slice
    in a `Molecule` / `Trajectory` pair costs 1.6 us each -- 0.55 s of the 1.2 s that configs[3] takes for its 325 000
    detected molecules, and a burst of garbage for the interpreter's collector -- whether or not the caller ever looks
    at them.  Indexing, slicing (a plain list), iteration, `len`, `==` with lists, `append` / `extend` behave like the
    list the reference returns; a molecule, once made, stays the same object.  Operations that rearrange the sequence
    (item assignment, deletion, `insert`, `pop`, `sort`, `reverse`, ...) first turn it into one plain list."""

    def __init__(self, molecules=()):
        self._chunks = []       # (rows, offsets, fates, names, strip) of one launch, or a plain list of Molecule objects
        self._starts = [0]      # index of every chunk's first molecule, and the total at the end
        self._made = {}
        self.extend(molecules)

    def add_rows(self, rows: np.ndarray, offsets, fates, names, strip_nans: bool = False) -> None:
        """Molecule k of this block owns rows[offsets[k]:offsets[k + 1]] and ended as names[fates[k]]."""
        if len(fates):
            self._chunks.append((rows, offsets, fates, names, bool(strip_nans)))
            self._starts.append(self._starts[-1] + len(fates))

    def extend(self, molecules) -> None:
        if isinstance(molecules, SavedMolecules):
            base = len(self)
            for chunk, lo, hi in zip(molecules._chunks, molecules._starts, molecules._starts[1:]):
                self._chunks.append(chunk)
                self._starts.append(self._starts[-1] + hi - lo)
            for i, m in molecules._made.items():
                self._made[base + i] = m
            return
        made = list(molecules)
        if made:
            self._chunks.append(made)
            self._starts.append(self._starts[-1] + len(made))

    def append(self, molecule) -> None:
        self.extend([molecule])

    def __len__(self) -> int:
        return self._starts[-1]

    def _make(self, i: int) -> "Molecule":
        c = bisect_right(self._starts, i) - 1
        chunk, k = self._chunks[c], i - self._starts[c]
        if isinstance(chunk, list):
            return chunk[k]
        rows, offsets, fates, names, strip = chunk
        name = names[fates[k]]
        m = Molecule.from_rows(rows[offsets[k]:offsets[k + 1]], name, name == "Detected")
        if strip:
            # Beamline.propagate_through ends with trajectory.drop_nans() (beamline.py:38, molecule.py:160-167), which
            # also strips rows that a non-finite initial condition or an overflow filled with NaN / inf
            m.trajectory.drop_nans()
            m.trajectory.n = m.trajectory.t.shape[0]
        return m

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        n = len(self)
        j = i + n if i < 0 else i
        if not 0 <= j < n:
            raise IndexError("molecule index out of range")
        m = self._made.get(j)
        if m is None:
            m = self._made[j] = self._make(j)
        return m

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    # -- rearranging: everything is made, then it is a list ---------------------------------------
    def _as_list(self) -> list:
        if not (len(self._chunks) == 1 and isinstance(self._chunks[0], list)):
            made = [self[i] for i in range(len(self))]
            self._chunks = [made] if made else []
            self._starts = [0, len(made)] if made else [0]
        self._made = {}
        return self._chunks[0] if self._chunks else []

    def _relist(self, made: list) -> None:
        self._chunks = [made] if made else []
        self._starts = [0, len(made)] if made else [0]
        self._made = {}

    def __setitem__(self, i, value) -> None:
        made = self._as_list()
        made[i] = value
        self._relist(made)

    def __delitem__(self, i) -> None:
        made = self._as_list()
        del made[i]
        self._relist(made)

    def insert(self, i, value) -> None:
        made = self._as_list()
        made.insert(i, value)
        self._relist(made)

    def sort(self, *, key=None, reverse=False) -> None:
        made = self._as_list()
        made.sort(key=key, reverse=reverse)
        self._relist(made)

    def __eq__(self, other):
        if isinstance(other, (list, tuple, SavedMolecules)):
            return len(self) == len(other) and all(a == b for a, b in zip(self, other))
        return NotImplemented

    __hash__ = None

    def __add__(self, other):
        out = SavedMolecules(self)
        out.extend(other)
        return out

    def __repr__(self) -> str:
        return f"SavedMolecules({len(self)} molecules, {len(self._made)} materialised)"
