"""Stand-in for h5py._hl (molecule.py:3 imports `group` from it)."""
group = None
