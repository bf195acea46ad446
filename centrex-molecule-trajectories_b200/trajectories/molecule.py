"""Molecule / Trajectory containers (reference molecule.py:8-186).

On the GPU path a molecule is nine FP64 registers; these classes are the host
view of a finished (or hand-built) trajectory: `x`, `v`, `a` are (n,3) float64
arrays, `t` is (n,), exactly what post-processing and the HDF layout expect.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

g = 9.80665  # scipy.constants.g


@dataclass
class Molecule:
    alive: bool = True

    def init_trajectory(self, beamline, x0=np.array((0, 0, 0)), v0=np.array((0, 0, 200)),
                        a0=np.array((0, -g, 0)), t0=0):
        self.trajectory = Trajectory(beamline)
        self.trajectory.update(x0, v0, a0, t0)

    # kinematics of the last row (molecule.py:26-56); `not delta_t` covers None and 0
    def x(self, delta_t: float = None):
        if not delta_t:
            return self.trajectory.x[self.trajectory.n - 1, :]
        return self.x() + self.v() * delta_t + self.a() * delta_t ** 2 / 2

    def v(self, delta_t: float = None):
        if not delta_t:
            return self.trajectory.v[self.trajectory.n - 1, :]
        return self.v() + self.a() * delta_t

    def a(self):
        return self.trajectory.a[self.trajectory.n - 1, :]

    def t(self):
        return self.trajectory.t[self.trajectory.n - 1]

    def update_trajectory(self, delta_t, a=np.array((0, -g, 0))):
        x, v, t = self.x(delta_t), self.v(delta_t), self.t() + delta_t
        self.trajectory.update(x, v, a, t)

    def set_aperture_hit(self, aperture_name):
        self.aperture_hit = aperture_name

    def set_dead(self):
        self.alive = False

    def plot_trajectory(self, axes):
        color = {"Detected": "g", "Field plates": "r"}.get(self.aperture_hit, "k")
        axes[0].plot(self.trajectory.x[:, 2], self.trajectory.x[:, 0], c=color)
        axes[1].plot(self.trajectory.x[:, 2], self.trajectory.x[:, 1], c=color)

    def save_to_hdf(self, file, run_name: str, group_name: str):
        self.trajectory.save_to_hdf(file, run_name, group_name)
        grp = file[run_name + "/" + group_name]
        grp.attrs["aperture_hit"] = self.aperture_hit
        grp.attrs["alive"] = self.alive

    @classmethod
    def from_rows(cls, rows: np.ndarray, aperture_hit: str, alive: bool) -> "Molecule":
        """Wrap GPU-produced rows [n,10] = x,y,z,vx,vy,vz,ax,ay,az,t."""
        m = cls(alive=alive)
        m.trajectory = Trajectory.from_rows(rows)
        m.aperture_hit = aperture_hit
        return m


class Trajectory:
    def __init__(self, beamline=None, n_rows: int = None):
        if n_rows is None:
            n_rows = 10 + sum(e.N_steps() for e in beamline.elements)   # molecule.py:117-121
        self.x = np.full((n_rows, 3), np.nan)
        self.v = np.full((n_rows, 3), np.nan)
        self.a = np.full((n_rows, 3), np.nan)
        self.t = np.full((n_rows,), np.nan)
        self.n = 0

    @classmethod
    def from_rows(cls, rows: np.ndarray) -> "Trajectory":
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        tr = cls.__new__(cls)
        tr.x, tr.v, tr.a, tr.t = rows[:, 0:3], rows[:, 3:6], rows[:, 6:9], rows[:, 9]
        tr.n = rows.shape[0]
        return tr

    def update(self, x, v, a, t):
        if self.n >= self.t.shape[0]:
            self._grow(max(16, self.n))
        self.x[self.n, :], self.v[self.n, :], self.a[self.n, :], self.t[self.n] = x, v, a, t
        self.n += 1

    def extend_rows(self, rows: np.ndarray):
        """Append GPU-produced rows [k,10]."""
        k = rows.shape[0]
        if self.n + k > self.t.shape[0]:
            self._grow(self.n + k - self.t.shape[0])
        s = slice(self.n, self.n + k)
        self.x[s], self.v[s], self.a[s], self.t[s] = rows[:, 0:3], rows[:, 3:6], rows[:, 6:9], rows[:, 9]
        self.n += k

    def _grow(self, extra: int):
        self.x = np.concatenate((self.x, np.full((extra, 3), np.nan)))
        self.v = np.concatenate((self.v, np.full((extra, 3), np.nan)))
        self.a = np.concatenate((self.a, np.full((extra, 3), np.nan)))
        self.t = np.concatenate((self.t, np.full((extra,), np.nan)))

    def add_steps(self, beamline):
        self._grow(sum(e.N_steps() for e in beamline.elements))

    def drop_nans(self):
        self.x = self.x[np.all(np.isfinite(self.x), axis=1), :]
        self.v = self.v[np.all(np.isfinite(self.v), axis=1), :]
        self.a = self.a[np.all(np.isfinite(self.a), axis=1), :]
        self.t = self.t[np.isfinite(self.t)]

    def save_to_hdf(self, file, run_name: str, group_name: str) -> None:
        self.drop_nans()
        path = run_name + "/" + group_name
        file.create_group(path)
        for key in ("x", "v", "a", "t"):
            file[path].create_dataset(key, data=getattr(self, key))
