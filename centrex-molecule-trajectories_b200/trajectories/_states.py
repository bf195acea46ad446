"""Minimal TlF basis-state bookkeeping used when `centrex_TlF` is not installed.

Only what the propagation path touches (electrostatic_lens.py:33-43,176-177 of
the reference): `amp * UncoupledBasisState(...)` builds a `State`, and
`State.find_largest_component()` gives back the component whose J, mJ select
the Stark curve.  Real picklable classes, so beamlines holding them can be
pickled or sent to other processes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Tuple


@dataclass(frozen=True)
class UncoupledBasisState:
    J: Any = 0
    mJ: Any = 0
    I1: Any = 0.5
    m1: Any = 0.5
    I2: Any = 0.5
    m2: Any = 0.5
    Omega: Any = 0
    P: Any = None
    electronic_state: Any = None

    def __rmul__(self, amp):
        return State([(amp, self)])

    __mul__ = __rmul__

    def __repr__(self) -> str:
        return (f"|{self.electronic_state}, J = {self.J}, mJ = {self.mJ}, I1 = {self.I1}, m1 = {self.m1}, "
                f"I2 = {self.I2}, m2 = {self.m2}, P = {self.P}, Omega = {self.Omega}>")


class State:
    def __init__(self, data=()):
        self.data: List[Tuple[Any, UncoupledBasisState]] = list(data)

    def find_largest_component(self) -> UncoupledBasisState:
        return max(self.data, key=lambda t: abs(t[0]))[1]

    def __rmul__(self, amp):
        return State([(amp * a, s) for a, s in self.data])

    def __add__(self, other):
        return State(self.data + other.data)

    def __bool__(self) -> bool:
        return True

    def __eq__(self, other) -> bool:
        return isinstance(other, State) and self.data == other.data

    def __repr__(self) -> str:
        return " + ".join(f"{a:.2f} x {s!r}" for a, s in self.data)
