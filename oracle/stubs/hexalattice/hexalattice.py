"""Stand-in for hexalattice.hexalattice (absent from this image; only the Honeycomb element uses it).

PARITY UNPINNED at this boundary: hexalattice is not vendored by the reference, not pinned by it
(README.md / setup.py name no version) and not installable here, so `make_grid` below RESTATES the
published algorithm of hexalattice 1.x (github.com/alexkaz2/hexalattice, `hexalattice.py: make_grid`)
from its documentation and source as the author of this repository knows it; it could not be executed
against the real package.  Call site in the reference: beamline_elements/meshes.py:54-62
(`nx, ny, n=0, min_diam, align_to_origin=True, crop_circ=0, rotate_deg=0`).

Layout: ny rows spaced min_diam*sqrt(3)/2, nx centres per row spaced min_diam, odd rows shifted by
min_diam/2, row-major order (row 0 first); with align_to_origin the centre of the "middle" hexagon
(column ceil(nx/2)-1, +1/2 when row ceil(ny/2)-1 is odd ... i.e. when ceil(ny/2) is even; row ceil(ny/2)-1)
is moved to the origin.  Returns two (nx*ny, 1) float64 arrays.
"""
import numpy as np


def make_grid(nx, ny, min_diam, n=0, crop_circ=0.0, rotate_deg=0.0, align_to_origin=True):
    ratio = np.sqrt(3) / 2
    if n > 0:
        ny = int(np.sqrt(n / ratio))
        nx = n // ny
    coord_x, coord_y = np.meshgrid(np.arange(nx), np.arange(ny), sparse=False, indexing="xy")
    coord_y = coord_y * ratio
    coord_x = coord_x.astype("float")
    coord_x[1::2, :] += 0.5
    coord_x = coord_x.reshape(-1, 1)
    coord_y = coord_y.reshape(-1, 1)
    coord_x *= min_diam
    coord_y = coord_y.astype("float") * min_diam
    mid_x = (np.ceil(nx / 2) - 1) + 0.5 * (np.ceil(ny / 2) % 2 == 0)
    mid_y = (np.ceil(ny / 2) - 1) * ratio
    mid_x *= min_diam
    mid_y *= min_diam
    if crop_circ > 0:
        rad = ((coord_x - mid_x) ** 2 + (coord_y - mid_y) ** 2) ** 0.5
        coord_x = coord_x[rad.flatten() <= crop_circ, :]
        coord_y = coord_y[rad.flatten() <= crop_circ, :]
    if not np.isclose(rotate_deg, 0):
        c, s = np.cos(np.deg2rad(rotate_deg)), np.sin(np.deg2rad(rotate_deg))
        rot = np.hstack((coord_x - mid_x, coord_y - mid_y)) @ np.array([[c, s], [-s, c]]).T
        coord_x, coord_y = np.hsplit(rot + np.array([mid_x, mid_y]), 2)
    if align_to_origin:
        coord_x -= mid_x
        coord_y -= mid_y
    return coord_x, coord_y
