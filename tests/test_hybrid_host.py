"""Host-side logic of mixed beamlines (built-in CUDA elements + user-defined BeamlineElement subclasses): which
elements run where.  The runs themselves need a GPU (tests/test_gpu_hybrid.py)."""
from tests.beamlines import lens_beamline, lens_table
from tests.test_gpu_hybrid import replaced, user_elements


def test_classification():
    from trajectories import _hybrid
    from trajectories.beamline_elements.apertures import CircularAperture

    PyCircular, BatchCircular, _ = user_elements()
    bl = lens_beamline(lens_table())
    assert not _hybrid.is_hybrid(bl.elements)
    assert all(_hybrid.runs_on_device(e) for e in bl.elements)
    assert not _hybrid.runs_on_device(PyCircular(name="p", z0=0.1, L=0.01))

    class Tweaked(CircularAperture):            # a subclass of a built-in type with its own stepping: host
        def propagate_through(self, molecule):
            pass

    class Renamed(CircularAperture):            # a subclass that only adds data: still the CUDA element
        pass

    assert not _hybrid.runs_on_device(Tweaked(name="t", z0=0.1, L=0.01))
    assert _hybrid.runs_on_device(Renamed(name="r", z0=0.1, L=0.01))
    assert _hybrid.is_hybrid(replaced(bl, "BB exit", PyCircular).elements)
