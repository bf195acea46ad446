"""HDF5 result layout of the reference.

Written through h5py when it is installed, otherwise through the package's own
dependency-free HDF5 writer/reader (`_minih5`, same interface, same file format),
so results can be saved and re-imported on a machine without libhdf5.

  /<run>/counter                        attrs {fate: count}            (trajectory_simulator.py:160-177)
  /<run>/beamline/<element.name>        attrs class + vars(element)    (apertures.py:56-80)
  /<run>/position_distribution          attrs class + fields           (distributions.py:122-142)
  /<run>/velocity_distribution          attrs class + fields           (distributions.py:78-98)
  /<run>/trajectories/molecule_<i>      datasets x,v,a,t + attrs aperture_hit, alive
"""
from __future__ import annotations


def h5py():
    """The HDF5 module to use: h5py if importable, else the built-in `_minih5`."""
    try:
        import h5py as _h5

        if _h5 is not None:
            return _h5
    except ImportError:
        pass
    from . import _minih5

    return _minih5


def save_element(element, filepath, parent_group_path: str) -> None:
    with h5py().File(filepath, "a") as f:
        try:
            path = parent_group_path + "/" + element.name
            f.create_group(path)
            f[path].attrs["class"] = type(element).__name__
            for key, value in vars(element).items():
                f[path].attrs[key] = value
        except ValueError:
            print("Can't save beamline element. Group already exists!")


def save_lens(lens, filepath, parent_group_path: str) -> None:
    # electrostatic_lens.py:145-166: a_interp skipped, state and falsy values stored as repr strings
    with h5py().File(filepath, "a") as f:
        path = parent_group_path + "/" + lens.name
        f.create_group(path)
        f[path].attrs["class"] = type(lens).__name__
        for key, value in vars(lens).items():
            if key == "a_interp":
                continue
            if key != "state" and value:
                f[path].attrs[key] = value
            else:
                f[path].attrs[key] = repr(value)


def save_beamline(beamline, filepath, run_name: str) -> None:
    group_path = run_name + "/beamline/"
    with h5py().File(filepath, "a") as f:
        f.create_group(group_path)
    for element in beamline.elements:
        element.save_to_hdf(filepath, group_path)


def save_distribution(dist, filepath, run_name: str, group: str) -> None:
    with h5py().File(filepath, "a") as f:
        try:
            path = run_name + "/" + group
            f.create_group(path)
            f[path].attrs["class"] = type(dist).__name__
            for key, value in vars(dist).items():
                f[path].attrs[key] = value
        except ValueError:
            raise ValueError(f"Can't save {group.replace('_', ' ')}. Group already exists!")


def save_counter(counter, filepath, run_name: str) -> None:
    with h5py().File(filepath, "a") as f:
        try:
            path = run_name + "/counter"
            f.create_group(path)
            for key, value in counter.counter_dict.items():
                f[path].attrs[key] = value
        except ValueError:
            raise ValueError("Can't save counter. Group already exists!")
