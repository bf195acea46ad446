"""Beamline elements with a CUDA implementation (reference: beamline_elements/__init__.py:1-3)."""
from .apertures import *  # noqa: F401,F403
from .apertures import BeamlineElement  # noqa: F401  (README-level name; the reference forgets to export it)
from .electrostatic_lens import *  # noqa: F401,F403
from .meshes import *  # noqa: F401,F403
