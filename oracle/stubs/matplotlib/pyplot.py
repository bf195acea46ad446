"""Stand-in for matplotlib.pyplot: every call is a no-op."""


def _noop(*args, **kwargs):
    return None


def subplots(*args, **kwargs):
    raise RuntimeError("matplotlib stub: plotting is not available in the oracle harness")


show = _noop
figure = _noop


class Axes:  # used in an annotation at def time (meshes.py:139)
    pass
