"""Honeycomb mesh element (API of the reference's beamline_elements/meshes.py:26-178).

A lattice of hexagonal cells in the xy plane: a molecule is assigned the cell whose centre is nearest
when it reaches z0 and must be inside that cell's polygon at z0 and at z1.  As everywhere in this package
the element carries geometry only; stepping and hit test run on the GPU (csrc/cmt_device.cuh: do_honeycomb).

The reference takes the cell centres from `hexalattice.make_grid` and the hit test from matplotlib's
`RegularPolygon.contains_point`.  Neither package is available to this build and the reference pins no
version of either, so both are RESTATED here from their published algorithms — parity is unpinned at that
third-party boundary; the reference's own logic around them (stepping, `if not idx`
re-assignment of cell 0, fate name) is pinned by tests/golden/honeycomb.npz.

Grid (hexalattice.make_grid(nx, ny, min_diam, n=0, crop_circ=0, rotate_deg=0, align_to_origin=True)):
ny rows spaced min_diam*sqrt(3)/2, nx centres per row spaced min_diam, odd rows shifted by min_diam/2,
numbered row by row; the centre of the middle cell is moved to the origin (x0, y0 do not enter).
"""
from __future__ import annotations

from dataclasses import dataclass
from math import ceil

import numpy as np

from .apertures import BeamlineElement

__all__ = ["Honeycomb"]


@dataclass
class Honeycomb(BeamlineElement):
    width: float = 2 * 25.4e-3
    height: float = 2 * 25.4e-3
    cell_wall_thickness: float = 1e-4
    cell_wall_length: float = 25.4e-3 / 16

    def __post_init__(self) -> None:
        super().__post_init__()
        self.x1 = self.x0 - self.width / 2
        self.x2 = self.x0 + self.width / 2
        self.y1 = self.y0 - self.height / 2
        self.y2 = self.y0 + self.height / 2
        # cells needed to cover width x height (meshes.py:50-51)
        self.nx = ceil(self.width / (self.cell_wall_length * np.sqrt(3)))
        self.ny = ceil(self.height / (self.cell_wall_length * 3 / 2))
        self.xcoords, self.ycoords = self.cell_centres()

    # -- geometry handed to the GPU ------------------------------------------------
    @property
    def pitch(self) -> float:
        """Centre spacing within a row: `min_diam` of make_grid (meshes.py:58)."""
        return float(self.cell_wall_length * np.sqrt(3))

    @property
    def polygon_radius(self) -> float:
        """Circum-radius of each cell's polygon (meshes.py:73-77)."""
        return float((self.cell_wall_length * np.sqrt(3) - self.cell_wall_thickness / 2) / 2)

    def grid_origin(self):
        """(mid_x, mid_y): position of the middle cell before the grid is aligned to the origin."""
        ratio = np.sqrt(3) / 2
        mid_x = (np.ceil(self.nx / 2) - 1) + 0.5 * (np.ceil(self.ny / 2) % 2 == 0)
        mid_y = (np.ceil(self.ny / 2) - 1) * ratio
        return float(mid_x * self.pitch), float(mid_y * self.pitch)

    def cell_centres(self):
        """Two (nx*ny, 1) arrays, the shape the reference keeps in `xcoords`, `ycoords`."""
        ratio = np.sqrt(3) / 2
        mid_x, mid_y = self.grid_origin()
        col, row = np.meshgrid(np.arange(self.nx, dtype=np.float64), np.arange(self.ny, dtype=np.float64))
        col[1::2, :] += 0.5
        xs = col.reshape(-1, 1) * self.pitch - mid_x
        ys = (row * ratio).reshape(-1, 1) * self.pitch - mid_y
        return xs, ys

    def N_steps(self) -> int:
        return 2

    def make_patches(self):
        """One matplotlib RegularPolygon per cell, as the reference keeps in `patches` (meshes.py:69-82; needs
        matplotlib).  The GPU path does not use them: its hit test restates `contains_point` (module docstring)."""
        from matplotlib.patches import RegularPolygon

        return [RegularPolygon((x, y), 6, radius=self.polygon_radius) for x, y in zip(self.xcoords, self.ycoords)]

    @property
    def patches(self):
        return self.make_patches()

    def plot_mesh(self, ax=None):
        """Draw the cells (needs matplotlib)."""
        import matplotlib.pyplot as plt
        from matplotlib.collections import PatchCollection
        from matplotlib.patches import RegularPolygon

        if ax is None:
            _, ax = plt.subplots()
        cells = [RegularPolygon((float(x), float(y)), 6, radius=self.polygon_radius)
                 for x, y in zip(self.xcoords[:, 0], self.ycoords[:, 0])]
        ax.add_collection(PatchCollection(cells))
        ax.set_title("Hexagonal mesh")
        ax.set_xlabel("X-position / m")
        ax.set_ylabel("Y-position / m")
        ax.set_xlim([-self.width / 2, self.width / 2])
        ax.set_ylim([-self.height / 2, self.height / 2])
        return ax
