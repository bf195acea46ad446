"""Stand-in for matplotlib.patches.

Rectangle is a name only.  RegularPolygon carries the one method the reference's hot path calls,
`contains_point` (beamline_elements/meshes.py:113), RESTATED from matplotlib's published algorithm —
PARITY UNPINNED at this boundary (matplotlib is absent from this image and the reference pins no version):

  * vertices: `Path.unit_regular_polygon(n)`: theta = 2*pi/n * arange(n + 1) + pi/2, (cos, sin), last vertex
    CLOSEPOLY (its coordinates are ignored, the sub-path closes on the first vertex);
  * patch transform: Affine2D().scale(radius).rotate(orientation).translate(x, y); with no Axes attached the
    Artist transform is the identity, so `contains_point` takes data coordinates;
  * picking radius: a default (filled, edgecolor "none") patch uses 0, i.e. the bare polygon;
  * `_path.point_in_path`: the crossings-multiply test of src/_path.h (`point_in_path_impl`):
    for every edge v0->v1 with (v0.y >= ty) != (v1.y >= ty), the point is toggled when
    ((v1.y - ty) * (v0.x - v1.x) >= (v1.x - tx) * (v0.y - v1.y)) == (v1.y >= ty); non-finite points are outside.
"""
import numpy as np


class _Patch:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


class Rectangle(_Patch):
    pass


class RegularPolygon(_Patch):
    def __init__(self, xy, numVertices, radius=5, orientation=0, **kwargs):
        super().__init__(xy, numVertices, radius=radius, orientation=orientation, **kwargs)
        self.xy, self.numvertices, self.radius, self.orientation = xy, numVertices, radius, orientation

    def _vertices(self):
        n = self.numvertices
        theta = (2 * np.pi / n) * np.arange(n + 1) + np.pi / 2.0
        ux, uy = np.cos(theta)[:n], np.sin(theta)[:n]
        # scale(r).rotate(o).translate(x, y) as one affine matrix, applied like agg::trans_affine::transform
        a, b = np.cos(self.orientation), np.sin(self.orientation)
        r = float(self.radius)
        sx, shx, shy, sy = a * r - b * 0.0, a * 0.0 - b * r, b * r + a * 0.0, b * 0.0 + a * r
        tx, ty = float(np.asarray(self.xy[0]).reshape(-1)[0]), float(np.asarray(self.xy[1]).reshape(-1)[0])
        return ux * sx + uy * shx + tx, ux * shy + uy * sy + ty

    def contains_point(self, point, radius=None):
        tx, ty = float(point[0]), float(point[1])
        if not (np.isfinite(tx) and np.isfinite(ty)):
            return False
        vx, vy = self._vertices()
        n = len(vx)
        inside = False
        for k in range(n):
            x0, y0, x1, y1 = vx[k], vy[k], vx[(k + 1) % n], vy[(k + 1) % n]
            f0, f1 = y0 >= ty, y1 >= ty
            if f0 != f1 and (((y1 - ty) * (x0 - x1) >= (x1 - tx) * (y0 - y1)) == f1):
                inside = not inside
        return bool(inside)
