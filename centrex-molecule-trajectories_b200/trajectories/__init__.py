"""`trajectories` -- drop-in for the Monte Carlo propagation path of
otimgren/centrex-molecule-trajectories, executed by hand-written sm_100a CUDA
kernels (libcmt_b200.so) instead of the reference's per-molecule Python loop.

Same import names as the reference package (src/trajectories/__init__.py:1-2),
so existing beamline scripts run unchanged:

    from trajectories.beamline import Beamline
    from trajectories.beamline_elements.apertures import CircularAperture, FieldPlates, RectangularAperture
    from trajectories.beamline_elements.electrostatic_lens import ElectrostaticLens
    from trajectories.trajectory_simulator import TrajectorySimulator
"""
from . import (beamline, beamline_elements, distributions, molecule, post_processing,  # noqa: F401
               stark_potential, trajectory_simulator, utils)

__all__ = ["beamline", "beamline_elements", "distributions", "molecule", "post_processing", "stark_potential",
           "trajectory_simulator", "utils"]
