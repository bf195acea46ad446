"""Optional stand-in for the external `centrex_TlF` package (github.com/ograsdijk/CeNTREX-TlF).

The reference's example `examples/lens_simulation_different_states.py` imports
`centrex_TlF.states.UncoupledBasisState` only to describe which rotational state
flies through the lens.  When the real package is not installed, put this
directory on PYTHONPATH *after* site-packages' would-be location:

    PYTHONPATH=centrex-molecule-trajectories_b200:centrex-molecule-trajectories_b200/shims python examples/...

It provides the state bookkeeping classes only; the Stark curve then comes from
the build's rigid-rotor model (`trajectories/_tlf.py`).
"""
