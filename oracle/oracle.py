"""ctypes front end of the CPU oracle (oracle/cmt_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.  The
shipped package (centrex-molecule-trajectories_b200/trajectories) never does.

Elements are read by duck typing (class name + the reference's attribute
names: apertures.py:22-36,83-91,147-163,213-225; electrostatic_lens.py:23-46),
so the same call works on reference objects and on the build's own classes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "libcmt_oracle.so"
G = 9.80665  # scipy.constants.g, molecule.py:6

CIRCULAR, RECTANGULAR, FIELDPLATES, LENS, HONEYCOMB = 0, 1, 2, 3, 4

ELEMENT_DTYPE = np.dtype(
    [
        ("type", "<i4"), ("fate", "<i4"), ("fate2", "<i4"), ("table", "<i4"),
        ("n_steps", "<i4"), ("pad_", "<i4"),
        ("z0", "<f8"), ("z1", "<f8"),
        ("x1", "<f8"), ("x2", "<f8"), ("y1", "<f8"), ("y2", "<f8"),
        ("R", "<f8"), ("dz", "<f8"),
    ],
    align=True,
)

SOURCE_DTYPE = np.dtype(
    [
        ("pos_kind", "<i4"), ("pad_", "<i4"),
        ("vmean", "<f8", (3,)), ("vsigma", "<f8", (3,)),
        ("p0", "<f8"), ("p1", "<f8"), ("z", "<f8"),
    ],
    align=True,
)

_lib = None


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc, no FMA contraction)."""
    src = HERE / "cmt_oracle.c"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-s"], check=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB_PATH))
        assert _lib.orc_sizeof_element() == ELEMENT_DTYPE.itemsize
        assert _lib.orc_sizeof_source() == SOURCE_DTYPE.itemsize
    return _lib


class Flat:
    """Flattened beamline: element table, lens tables, fate names."""

    def __init__(self, elements, fate_names, tab_r, tab_a, tab_off, tab_len, max_rows):
        self.elements = elements
        self.fate_names = fate_names
        self.tab_r, self.tab_a, self.tab_off, self.tab_len = tab_r, tab_a, tab_off, tab_len
        self.max_rows = max_rows

    @property
    def fate_detected(self) -> int:
        return self.fate_names.index("Detected")


def flatten(elements) -> Flat:
    """Element list (sorted by z0 like Beamline.__post_init__, beamline.py:17-18,40-45)."""
    elements = sorted(elements, key=lambda e: e.z0)
    names: list[str] = []

    def fid(name: str) -> int:
        if name not in names:
            names.append(name)
        return names.index(name)

    tab = np.zeros(len(elements), dtype=ELEMENT_DTYPE)
    rs, as_, off, ln = [], [], [], []
    max_rows = 1
    for i, e in enumerate(elements):
        kind = type(e).__name__
        t = tab[i]
        t["z0"], t["z1"] = e.z0, e.z1
        if kind == "CircularAperture":
            t["type"], t["fate"], t["R"] = CIRCULAR, fid(e.name), e.d / 2
            max_rows += 2
        elif kind == "RectangularAperture":
            t["type"], t["fate"] = RECTANGULAR, fid(e.name)
            t["x1"], t["x2"], t["y1"], t["y2"] = e.x1, e.x2, e.y1, e.y2
            max_rows += 2
        elif kind == "FieldPlates":
            t["type"], t["fate"] = FIELDPLATES, fid(e.name)
            t["x1"], t["x2"] = e.x1, e.x2
            max_rows += 2
        elif kind == "ElectrostaticLens":
            t["type"] = LENS
            t["fate"], t["fate2"] = fid("Lens entrance"), fid("Inside lens")
            t["R"], t["dz"] = e.d / 2, e.dz
            t["n_steps"] = int(np.rint(e.L / e.dz))
            if e.a_interp is None:
                raise ValueError("oracle: the lens acceleration table must be injected (a_interp)")
            x = np.ascontiguousarray(np.asarray(e.a_interp.x, dtype=np.float64))
            y = np.ascontiguousarray(np.asarray(e.a_interp.y, dtype=np.float64))
            t["table"] = len(off)
            off.append(sum(ln))
            ln.append(len(x))
            rs.append(x)
            as_.append(y)
            max_rows += 2 + int(t["n_steps"])
        elif kind == "Honeycomb":
            # meshes.py:50-62,73-77 + hexalattice.make_grid's origin (restated, see cmt_oracle.c header)
            nx = int(np.ceil(e.width / (e.cell_wall_length * np.sqrt(3))))
            ny = int(np.ceil(e.height / (e.cell_wall_length * 3 / 2)))
            pitch = e.cell_wall_length * np.sqrt(3)
            mid_x = (np.ceil(nx / 2) - 1) + 0.5 * (np.ceil(ny / 2) % 2 == 0)
            mid_y = (np.ceil(ny / 2) - 1) * (np.sqrt(3) / 2)
            t["type"], t["fate"] = HONEYCOMB, fid(e.name)
            t["n_steps"], t["pad_"] = nx, ny
            t["dz"], t["x1"], t["y1"] = pitch, mid_x * pitch, mid_y * pitch
            t["R"] = (e.cell_wall_length * np.sqrt(3) - e.cell_wall_thickness / 2) / 2
            max_rows += 2
        else:
            raise TypeError(f"oracle: unsupported beamline element {kind}")
    fid("Detected")
    tab_r = np.concatenate(rs) if rs else np.zeros(1)
    tab_a = np.concatenate(as_) if as_ else np.zeros(1)
    return Flat(tab, names, tab_r, tab_a, np.asarray(off or [0], dtype=np.int32),
                np.asarray(ln or [0], dtype=np.int32), max_rows)


def _p(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


def propagate(elements, ic, want_rows: bool = False, n_threads: int = 0):
    """Propagate ICs (6,n) through the beamline; returns a dict of per-molecule results.

    fate (n) int32 index into fate_names; fin (10,n) = x,y,z,vx,vy,vz,ax,ay,az,t of the
    last trajectory row; n_rows (n); rows (n,max_rows,10) NaN padded (optional);
    counters (n_fates) int64; work = [ballistic rows, lens steps, table out-of-range].
    """
    flat = elements if isinstance(elements, Flat) else flatten(elements)
    ic = np.ascontiguousarray(ic, dtype=np.float64)
    assert ic.ndim == 2 and ic.shape[0] == 6
    n = ic.shape[1]
    fate = np.empty(n, dtype=np.int32)
    fin = np.empty((10, n), dtype=np.float64)
    n_rows = np.empty(n, dtype=np.int32)
    rows = np.full((n, flat.max_rows, 10), np.nan) if want_rows else None
    counters = np.zeros(len(flat.fate_names), dtype=np.int64)
    work = np.zeros(3, dtype=np.int64)
    rc = lib().orc_propagate(
        flat.elements.ctypes.data_as(C.c_void_p), C.c_int(len(flat.elements)),
        _p(flat.tab_r, C.c_double), _p(flat.tab_a, C.c_double),
        _p(flat.tab_off, C.c_int32), _p(flat.tab_len, C.c_int32),
        C.c_int(flat.fate_detected), C.c_double(G),
        C.c_long(n), _p(ic, C.c_double), _p(fate, C.c_int32), _p(fin, C.c_double),
        _p(n_rows, C.c_int32), _p(rows, C.c_double), C.c_long(flat.max_rows),
        _p(counters, C.c_int64), C.c_int(len(counters)), _p(work, C.c_int64),
        C.c_int(n_threads),
    )
    assert rc == 0
    return dict(fate=fate, fin=fin, n_rows=n_rows, rows=rows, counters=counters, work=work,
                fate_names=flat.fate_names, flat=flat)


def make_source(vdist, xdist) -> np.ndarray:
    """Distribution objects (distributions.py:54-76,101-119,144-162) -> source record."""
    s = np.zeros(1, dtype=SOURCE_DTYPE)
    s["vmean"][0] = (vdist.vx, vdist.vy, vdist.vz)
    s["vsigma"][0] = (vdist.sigmax, vdist.sigmay, vdist.sigmaz)
    kind = type(xdist).__name__
    if kind == "CeNTREXPositionDistribution":
        s["pos_kind"], s["p0"], s["p1"] = 0, xdist.d / 2, 0.0
    elif kind == "GaussianPositionDistribution":
        s["pos_kind"], s["p0"], s["p1"] = 1, xdist.sigmax, xdist.sigmay
    else:
        raise TypeError(f"oracle: unsupported position distribution {kind}")
    s["z"] = xdist.z
    return s


def draw(source: np.ndarray, seed: int, first: int, n: int, n_threads: int = 0) -> np.ndarray:
    ic = np.empty((6, n), dtype=np.float64)
    lib().orc_draw(source.ctypes.data_as(C.c_void_p), C.c_uint64(seed), C.c_uint64(first),
                   C.c_long(n), _p(ic, C.c_double), C.c_int(n_threads))
    return ic


def philox4x32_10(ctr, key) -> np.ndarray:
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(ctr, C.c_uint32), _p(key, C.c_uint32), _p(out, C.c_uint32))
    return out


def run(elements, source: np.ndarray, seed: int, first: int, n: int, n_threads: int = 0):
    """Draw + propagate + count on the host cores (the timed CPU baseline)."""
    flat = elements if isinstance(elements, Flat) else flatten(elements)
    counters = np.zeros(len(flat.fate_names), dtype=np.int64)
    work = np.zeros(3, dtype=np.int64)
    rc = lib().orc_run(
        flat.elements.ctypes.data_as(C.c_void_p), C.c_int(len(flat.elements)),
        _p(flat.tab_r, C.c_double), _p(flat.tab_a, C.c_double),
        _p(flat.tab_off, C.c_int32), _p(flat.tab_len, C.c_int32),
        C.c_int(flat.fate_detected), C.c_double(G),
        source.ctypes.data_as(C.c_void_p), C.c_uint64(seed), C.c_uint64(first), C.c_long(n),
        _p(counters, C.c_int64), C.c_int(len(counters)), _p(work, C.c_int64), C.c_int(n_threads),
    )
    assert rc == 0
    return dict(counters=counters, work=work, fate_names=flat.fate_names)


def max_threads() -> int:
    return int(lib().orc_max_threads())


def host_cores() -> int:
    """Cores this process may run on.  torchrun exports OMP_NUM_THREADS=1, so callers that
    want every core pass this explicitly as n_threads."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1
