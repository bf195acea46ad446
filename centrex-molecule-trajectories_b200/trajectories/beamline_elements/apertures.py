"""Passive beamline elements: circular / rectangular apertures and field plates.

API mirror of the reference's `beamline_elements/apertures.py` (same class
names, dataclass fields, defaults and derived attributes), but the elements
carry geometry only: the stepping and hit tests run on the GPU
(csrc/cmt_device.cuh: do_circular, do_rectangular, do_fieldplates).
`propagate_through(molecule)` is kept as the plugin entry point of the
reference (apertures.py:38-42) and runs the same CUDA code on that one molecule.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from pathlib import Path

__all__ = ["CircularAperture", "RectangularAperture", "FieldPlates"]


@dataclass
class BeamlineElement(ABC):
    """Base of all beamline elements (reference apertures.py:22-54)."""

    name: str          # fate string recorded for molecules that hit the element
    z0: float          # z where the element starts / m
    L: float           # extent along z / m
    x0: float = 0.0    # centre of the element in x / m
    y0: float = 0.0    # centre of the element in y / m

    def __post_init__(self):
        self.z1 = self.z0 + self.L

    # -- plugin API ---------------------------------------------------------
    def propagate_through(self, molecule) -> None:
        """Advance `molecule` through this element on the GPU, appending the same
        trajectory rows and setting `alive`/`aperture_hit` as the reference does."""
        from .._single import propagate_molecule

        propagate_molecule([self], molecule, mark_detected=False)

    @abstractmethod
    def N_steps(self) -> int:
        """Upper bound of trajectory rows this element appends."""

    def plot(self, axes) -> None:
        """Draw the element on (XZ, YZ) axes; needs matplotlib."""
        from .._plotting import plot_element

        plot_element(self, axes)

    def save_to_hdf(self, filepath: Path, parent_group_path: str) -> None:
        """attrs `class` + every instance attribute on group <parent>/<name> (apertures.py:56-80)."""
        from .._hdf import save_element

        save_element(self, filepath, parent_group_path)


@dataclass
class CircularAperture(BeamlineElement):
    """Circular opening of diameter d, tested at z0 and z1 about the beam axis
    (x0, y0 are ignored by the test, as in apertures.py:110-111)."""

    d: float = 0.0254

    def N_steps(self) -> int:
        return 2


@dataclass
class RectangularAperture(BeamlineElement):
    """Rectangular opening: w spans x, h spans y, centred on (x0, y0); strict
    inequalities at z0 and z1 (apertures.py:157-163,183-186)."""

    w: float = 0.02
    h: float = 0.02

    def __post_init__(self):
        super().__post_init__()
        self.x1 = self.x0 - self.w / 2
        self.x2 = self.x0 + self.w / 2
        self.y1 = self.y0 - self.h / 2
        self.y2 = self.y0 + self.h / 2

    def N_steps(self) -> int:
        return 2


@dataclass
class FieldPlates(BeamlineElement):
    """Parallel plates bounding x over [z0, z1]; a molecule that would leave the
    gap before z1 stops where it crosses the plate (apertures.py:227-270)."""

    w: float = 0.02

    def __post_init__(self):
        super().__post_init__()
        self.x1 = self.x0 - self.w / 2
        self.x2 = self.x0 + self.w / 2

    def N_steps(self) -> int:
        return 2
