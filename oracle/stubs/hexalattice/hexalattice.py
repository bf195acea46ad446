def make_grid(*args, **kwargs):
    raise RuntimeError("hexalattice stub: Honeycomb is outside the oracle harness")
