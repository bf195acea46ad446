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


def propagate_molecule(elements, molecule, mark_detected: bool) -> None:
    torch = eng._torch()
    if not getattr(molecule, "alive", True):
        return
    prop = eng.Propagator(eng.flatten(sorted(elements, key=lambda e: e.z0)))
    tr = molecule.trajectory
    last = tr.n - 1
    state = np.empty((10, 1), dtype=np.float64)
    state[0:3, 0], state[3:6, 0], state[6:9, 0], state[9, 0] = tr.x[last], tr.v[last], tr.a[last], tr.t[last]
    if state[8, 0] != 0.0:
        raise ValueError("a_z != 0 is not supported on the GPU path")
    rows, offsets, fate = prop.trajectories(torch.from_numpy(state).to(prop.tdev))
    tr.extend_rows(rows[1:int(offsets[1])])   # row 0 repeats the molecule's current row
    name = prop.flat.fate_names[int(fate[0])]
    if name == "Detected":
        if mark_detected:
            molecule.set_aperture_hit("Detected")
    else:
        molecule.set_dead()
        molecule.set_aperture_hit(name)
    if mark_detected:
        tr.drop_nans()
