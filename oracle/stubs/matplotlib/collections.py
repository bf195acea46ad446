"""Stand-in for matplotlib.collections (names only)."""


class PatchCollection:
    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs
