"""Honeycomb mesh element: declared for import compatibility, not implemented on the GPU.

The reference's Honeycomb (meshes.py:26-178) takes its cell centres from
`hexalattice.make_grid` and its hit test from matplotlib's `RegularPolygon.contains_point`;
neither package is available to this build, so there is nothing to pin a CUDA
implementation to, and no example beamline uses the element.  The class keeps the
constructor signature so that scripts importing it still load; putting one into a
Beamline that is then simulated raises TypeError from the flattening step.
"""
from __future__ import annotations

from dataclasses import dataclass

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

    def N_steps(self) -> int:
        return 2

    def propagate_through(self, molecule) -> None:
        raise NotImplementedError(
            "Honeycomb has no CUDA implementation in this build (its geometry and hit test come from "
            "hexalattice and matplotlib in the reference, which are not available to pin parity to)")
