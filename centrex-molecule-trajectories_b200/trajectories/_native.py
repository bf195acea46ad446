"""ctypes binding of libcmt_b200.so (the C ABI declared in include/cmt.h).

There is no CPU fallback: if the CUDA library has not been built, or no B200 is
visible, every compute entry point raises.  Build it with
`python -c "import __graft_entry__ as g; g.build()"` from the repository root
(or `make -C centrex-molecule-trajectories_b200`).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

PKG_ROOT = Path(__file__).resolve().parent.parent          # centrex-molecule-trajectories_b200/
LIB_PATH = PKG_ROOT / "lib" / "libcmt_b200.so"

CMT_MAX_ELEMENTS = 40
CMT_MAX_FATES = 64
CMT_MAX_TABLES = 8
CMT_ROW_DOUBLES = 10
CMT_MAX_PLANES = 16
CIRCULAR, RECTANGULAR, FIELDPLATES, LENS, HONEYCOMB = 0, 1, 2, 3, 4
POS_DISC, POS_GAUSS = 0, 1
MATH_EXACT, MATH_CONTRACTED = 0, 1
MATH_MODES = {"exact": MATH_EXACT, "contracted": MATH_CONTRACTED}


class Element(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("fate", C.c_int32), ("fate2", C.c_int32), ("table", C.c_int32),
        ("n_steps", C.c_int32), ("reserved", C.c_int32),
        ("z0", C.c_double), ("z1", C.c_double),
        ("x1", C.c_double), ("x2", C.c_double),
        ("y1", C.c_double), ("y2", C.c_double),
        ("R", C.c_double), ("dz", C.c_double),
    ]


class Table(C.Structure):
    _fields_ = [("r", C.POINTER(C.c_double)), ("a", C.POINTER(C.c_double)),
                ("n", C.c_int32), ("reserved", C.c_int32)]


class Source(C.Structure):
    _fields_ = [
        ("pos_kind", C.c_int32), ("reserved", C.c_int32),
        ("vmean", C.c_double * 3), ("vsigma", C.c_double * 3),
        ("p0", C.c_double), ("p1", C.c_double), ("z", C.c_double),
    ]


class Outputs(C.Structure):
    _fields_ = [
        ("fate", C.c_void_p), ("final_state", C.c_void_p), ("final_ld", C.c_int64),
        ("counters", C.c_void_p), ("work", C.c_void_p),
        ("saved_index", C.c_void_p), ("saved_count", C.c_void_p),
        ("saved_capacity", C.c_int64), ("save_mask", C.c_uint64),
        ("queue_capacity", C.c_int64),
    ]


class NativeError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    "cmt_beamline_create": (C.c_int, [C.POINTER(Element), C.c_int, C.POINTER(Table), C.c_int, C.c_int,
                                      C.c_int, C.c_double, C.c_int, C.POINTER(C.c_void_p)]),
    "cmt_beamline_destroy": (None, [C.c_void_p]),
    "cmt_beamline_set_math": (C.c_int, [C.c_void_p, C.c_int]),
    "cmt_beamline_max_rows": (C.c_int, [C.c_void_p]),
    "cmt_beamline_device": (C.c_int, [C.c_void_p]),
    "cmt_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64]),
    "cmt_propagate_ic": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                   C.POINTER(Outputs), C.c_void_p, C.c_size_t, C.c_void_p]),
    "cmt_propagate_philox": (C.c_int, [C.c_void_p, C.POINTER(Source), C.c_uint64, C.c_int64, C.c_int64,
                                       C.POINTER(Outputs), C.c_void_p, C.c_size_t, C.c_void_p]),
    "cmt_philox_draw": (C.c_int, [C.POINTER(Source), C.c_uint64, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_int64, C.c_void_p]),
    "cmt_trajectories": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                   C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "cmt_resume": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
                             C.c_void_p, C.c_void_p]),
    "cmt_plane_crossings": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                      C.c_int64, C.POINTER(C.c_double), C.c_int32, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmt_run_host_ic": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "cmt_run_host_philox": (C.c_int, [C.c_void_p, C.POINTER(Source), C.c_uint64, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_void_p]),
    "cmt_timing_enable": (C.c_int, [C.c_int]),
    "cmt_timing_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "cmt_timing_timeline": (C.c_int64, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32), C.c_int64]),
    "cmt_launch_count": (C.c_int64, [C.c_int]),
    "cmt_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cmt_selftest": (C.c_int, [C.c_int, C.c_int64, C.c_uint64, C.c_int, C.POINTER(C.c_int64)]),
    "cmt_debug_flags": (C.c_int, [C.c_int]),
    "cmt_version": (C.c_int, []),
    "cmt_last_error": (C.c_char_p, []),
}

EXPORTS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    """Load libcmt_b200.so; raises NativeError when it has not been built."""
    global _lib
    if _lib is None:
        path = Path(os.environ.get("CMT_B200_LIB", LIB_PATH))
        if not path.exists():
            raise NativeError(
                f"{path} not found: the CUDA library is not built and there is no CPU fallback. "
                "Run __graft_entry__.build() (nvcc, sm_100a)."
            )
        handle = C.CDLL(str(path))
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().cmt_last_error().decode("utf-8", "replace")
        raise NativeError(f"libcmt_b200 error {rc}: {msg}")
