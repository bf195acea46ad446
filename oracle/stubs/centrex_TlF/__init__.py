"""Import-time stand-in for centrex_TlF (external, unpinned, absent).

Only the state bookkeeping classes the reference touches at import time and on
the lens path (electrostatic_lens.py:10,33-43,176-177) are provided, as real
picklable classes so joblib/loky workers can un-pickle a beamline.  The Stark
Hamiltonian itself is NOT provided: parity at that boundary is unpinned and the
lens acceleration table is always injected through `ElectrostaticLens.a_interp`.
"""
