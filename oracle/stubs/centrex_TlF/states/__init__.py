from dataclasses import dataclass
from typing import Any


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


class State:
    def __init__(self, data=()):
        self.data = list(data)

    def find_largest_component(self):
        return max(self.data, key=lambda t: abs(t[0]))[1]

    def state_vector(self, QN):
        raise RuntimeError("centrex_TlF stub: no Hamiltonian available")

    def __rmul__(self, amp):
        return State([(amp * a, s) for a, s in self.data])

    def __repr__(self):
        return " + ".join(f"{a} x {s!r}" for a, s in self.data)

    def __bool__(self):
        return True


def generate_uncoupled_states_ground(Js):
    raise RuntimeError("centrex_TlF stub: no Hamiltonian available")
