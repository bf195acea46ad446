"""Single-molecule entry points of the plugin API, executed on the GPU.

`BeamlineElement.propagate_through(molecule)` (apertures.py:38-42) and
`Beamline.propagate_through(molecule)` (beamline.py:20-38) mutate one Molecule:
append trajectory rows, set `alive`/`aperture_hit`.  Here the molecule's last
row (x, v, a, t) is sent to the trajectory kernel, which resumes from it and
returns every new row.
"""
from __future__ import annotations

import numpy as np

from . import _engine as eng


def _device_run(elements, molecule) -> None:
    """Built-in elements: the molecule's last row goes to the trajectory kernel, every new row comes back."""
    torch = eng._torch()
    prop = eng.Propagator(eng.flatten(elements))
    tr = molecule.trajectory
    last = tr.n - 1
    state = np.empty((10, 1), dtype=np.float64)
    state[0:3, 0], state[3:6, 0], state[6:9, 0], state[9, 0] = tr.x[last], tr.v[last], tr.a[last], tr.t[last]
    if state[8, 0] != 0.0:
        raise ValueError("a_z != 0 is not supported on the GPU path")
    rows, offsets, fate = prop.trajectories(torch.from_numpy(state).to(prop.tdev))
    tr.extend_rows(rows[1:int(offsets[1])])   # row 0 repeats the molecule's current row
    name = prop.flat.fate_names[int(fate[0])]
    if name != "Detected":
        molecule.set_dead()
        molecule.set_aperture_hit(name)


def lens_interior(lens, molecule) -> None:
    """ElectrostaticLens.propagate_inside_lens: the device run of the lens alone from the molecule's last row, minus the
    row at the entrance plane and the exit row that ElectrostaticLens.propagate_through puts around the integration."""
    torch = eng._torch()
    if not getattr(molecule, "alive", True):
        return
    prop = eng.Propagator(eng.flatten([lens]))
    tr = molecule.trajectory
    last = tr.n - 1
    state = np.empty((10, 1), dtype=np.float64)
    state[0:3, 0], state[3:6, 0], state[6:9, 0], state[9, 0] = tr.x[last], tr.v[last], tr.a[last], tr.t[last]
    if state[8, 0] != 0.0:
        raise ValueError("a_z != 0 is not supported on the GPU path")
    rows, offsets, fate = prop.trajectories(torch.from_numpy(state).to(prop.tdev))
    rows = rows[1:int(offsets[1])]                   # row 0 repeats the molecule's current row
    name = prop.flat.fate_names[int(fate[0])]
    if name == "Lens entrance":
        tr.extend_rows(rows)
    elif name == "Inside lens":
        tr.extend_rows(rows[1:])                     # without the row at z0
    else:
        tr.extend_rows(rows[1:-1])                   # without the row at z0 and the exit row at z1
    if name in ("Lens entrance", "Inside lens"):
        molecule.set_dead()
        molecule.set_aperture_hit(name)


def propagate_molecule(elements, molecule, mark_detected: bool) -> None:
    from ._hybrid import runs_on_device

    if not getattr(molecule, "alive", True):
        return
    run = []
    for element in sorted(elements, key=lambda e: e.z0):
        if runs_on_device(element):
            run.append(element)
            continue
        if run:
            _device_run(run, molecule)
            run = []
        if not molecule.alive:
            break
        element.propagate_through(molecule)          # a user-defined element: its own Python code (beamline.py:26-31)
        if not molecule.alive:
            break
    if run and molecule.alive:
        _device_run(run, molecule)
    if mark_detected:
        if molecule.alive:
            molecule.set_aperture_hit("Detected")
        molecule.trajectory.drop_nans()
