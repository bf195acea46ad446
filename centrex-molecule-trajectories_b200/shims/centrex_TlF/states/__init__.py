from trajectories._states import State, UncoupledBasisState  # noqa: F401
