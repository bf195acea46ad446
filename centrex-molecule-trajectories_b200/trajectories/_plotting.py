"""matplotlib drawings of beamline elements (optional dependency, imported lazily).

Out of the hot-path scope; kept so scripts that call Beamline.plot() keep working
when matplotlib is installed (reference beamline.py:57-76 and the elements' plot()).
"""
from __future__ import annotations


def _patches():
    try:
        from matplotlib.patches import Rectangle
    except ImportError as e:  # pragma: no cover
        raise ImportError("plotting needs matplotlib, which is not installed") from e
    return Rectangle


def plot_element(e, axes) -> None:
    Rectangle = _patches()
    kind = type(e).__name__
    L = e.z1 - e.z0
    if kind == "CircularAperture":
        for ax in axes[:2]:
            ax.add_patch(Rectangle((e.z0, e.d / 2), L, 1, color=(0.5, 0.5, 0.5)))
            ax.add_patch(Rectangle((e.z0, -e.d / 2 - 1), L, 1, color=(0.5, 0.5, 0.5)))
    elif kind == "RectangularAperture":
        axes[0].add_patch(Rectangle((e.z0, e.x2), L, 0.05, color="k"))
        axes[0].add_patch(Rectangle((e.z0, e.x1 - 0.05), L, 0.05, color="k"))
        axes[1].add_patch(Rectangle((e.z0, e.y2), L, 0.05, color="k"))
        axes[1].add_patch(Rectangle((e.z0, e.y1 - 0.05), L, 0.05, color="k"))
    elif kind == "FieldPlates":
        axes[0].add_patch(Rectangle((e.z0, e.x2), L, 0.02, color="y"))
        axes[0].add_patch(Rectangle((e.z0, e.x1 - 0.02), L, 0.02, color="y"))
    elif kind == "Honeycomb":          # meshes.py:126-138
        axes[0].add_patch(Rectangle((e.z0, e.x1), L, e.x2 - e.x1, color="k"))
        axes[1].add_patch(Rectangle((e.z0, e.y1), L, e.y2 - e.y1, color="k"))
    elif kind == "ElectrostaticLens":
        for ax in axes[:2]:
            ax.add_patch(Rectangle((e.z0, e.d / 2), L, 0.02, color="b"))
            ax.add_patch(Rectangle((e.z0, -e.d / 2 - 0.02), L, 0.02, color="b"))


def plot_beamline(beamline):
    try:
        import matplotlib.pyplot as plt
    except ImportError as e:  # pragma: no cover
        raise ImportError("plotting needs matplotlib, which is not installed") from e
    fig, axes = plt.subplots(2, 1, figsize=(16, 9))
    zmax = beamline.elements[-1].z1 + 0.1
    for ax, label in zip(axes, ("X-position / m", "Y-position / m")):
        ax.set_ylabel(label)
        ax.set_ylim([-0.06, 0.06])
        ax.set_xlim([0, zmax])
    axes[1].set_xlabel("Z-position / m")
    for element in beamline.elements:
        element.plot(axes)
    return axes
