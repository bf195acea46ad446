"""Beamline: an ordered collection of elements (reference beamline.py:10-91)."""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import List


@dataclass
class Beamline:
    elements: List

    def __post_init__(self):
        self.sort_elements()

    def sort_elements(self) -> None:
        """Sort the caller's list in place by z0 (flight order), as beamline.py:40-45 does."""
        self.elements.sort(key=lambda e: e.z0)

    def propagate_through(self, molecule) -> None:
        """Fly one molecule through every element on the GPU (beamline.py:20-38):
        rows are appended to its trajectory, a hit sets alive=False and
        aperture_hit, a survivor is marked "Detected"."""
        from ._single import propagate_molecule

        propagate_molecule(self.elements, molecule, mark_detected=True)

    def find_element(self, name):
        for element in self.elements:
            if element.name == name:
                return element
        print(f"Element with name '{name}' not found in beamline")

    def plot(self):
        from ._plotting import plot_beamline

        return plot_beamline(self)

    def save_to_hdf(self, filepath: Path, run_name: str) -> None:
        from ._hdf import save_beamline

        save_beamline(self, filepath, run_name)
