"""Stand-in for matplotlib.patches (names only)."""


class _Patch:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


class Rectangle(_Patch):
    pass


class RegularPolygon(_Patch):
    pass
