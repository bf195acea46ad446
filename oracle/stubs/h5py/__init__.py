"""Import-time stand-in for h5py (absent in this image; no libhdf5).

The reference uses `h5py.File` in annotations evaluated at def time
(molecule.py:97,169), so the name must exist.
"""


class File:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("h5py stub: HDF5 is not available in the oracle harness")
