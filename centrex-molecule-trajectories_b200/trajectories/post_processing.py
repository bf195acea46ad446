"""Distributions of saved molecules at a plane z (reference post_processing.py:20-140).

Host-side analysis of the trajectories the GPU path hands back: for every saved
molecule that reached z, take the last stored row before the plane and fly
ballistically with that row's acceleration (or return the stored row when z is
one of its rows).  Same results and shapes as the reference; one shared helper
instead of two copies of the loop.

The same quantities without any stored trajectory (evaluated inside the
propagation kernel, 40 B per molecule and plane): `TrajectorySimulator.plane_distributions`.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

__all__ = ["take_timestep", "find_radial_pos_dist", "find_vel_dist", "state_at_plane"]


def take_timestep(x0: np.ndarray, v0: np.ndarray, a0: np.ndarray, dt: float):
    """Position and velocity after dt under constant acceleration (post_processing.py:9-17)."""
    return x0 + v0 * dt + a0 * dt ** 2 / 2, v0 + a0 * dt


def state_at_plane(molecule, z: float):
    """(x, v) of one molecule at the plane z, or None when it never got there."""
    tr = molecule.trajectory
    zs = tr.x[: tr.n, 2] if hasattr(tr, "n") else tr.x[:, 2]
    if zs[-1] < z:
        return None
    exact = np.nonzero(zs == z)[0]
    if exact.size:
        k = int(exact[0])
        return tr.x[k], tr.v[k]
    k = int(np.argmax(~(zs < z))) - 1          # last row strictly before the plane
    x, v, a = tr.x[k], tr.v[k], tr.a[k]
    return take_timestep(x, v, a, (z - x[2]) / v[2])


def _collect(result, z: float, elements: Optional[List[str]], pick):
    out = []
    for molecule in result.molecules:
        if elements is not None and molecule.aperture_hit not in elements:
            continue
        state = state_at_plane(molecule, z)
        if state is not None:
            out.append(pick(*state))
    return np.array(out)


def find_radial_pos_dist(result, z: float, elements: List[str] = None) -> np.ndarray:
    """(n, 2) array of x, y at the plane z for the saved molecules that reached it."""
    return _collect(result, z, elements, lambda x, v: x[:2])


def find_vel_dist(result, z: float, elements: List[str] = None) -> np.ndarray:
    """(n, 3) array of velocities at the plane z for the saved molecules that reached it."""
    return _collect(result, z, elements, lambda x, v: v)


def plot_radial_pos_dist(result, z: float):
    """Not implemented in the reference either (post_processing.py:143-149)."""


def plot_radial_vel_dist(result, z: float):
    """Not implemented in the reference either (post_processing.py:152-158)."""
