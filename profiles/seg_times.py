#!/usr/bin/env python
"""Duration of each lens segment kernel of a LONE step (one stream), for a list of launch sizes: what a segment costs
as a function of how many groups of 32 molecules it has per warp slot (CMT_TUNE_SEG_CTAS sets the CTAs per SM).

    CMT_TUNE_SEG_CTAS=3 python profiles/seg_times.py --molecules 2.5e6,5e6,1e7,2e7 [--reps 10]

One JSON line per size: molecules, groups per segment (from the work counters of one step), median duration of each
segment kernel and of the whole lens stage in ms (CUDA events on the launch stream, cmt_timing_enable(2))."""
import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from trajectories import _engine as eng  # noqa: E402
from trajectories import _native as nat  # noqa: E402
from trajectories.centrex import lens_beamline, lens_table  # noqa: E402
from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution  # noqa: E402


def timeline(lib, cap=1 << 16):
    a, b = (C.c_double * cap)(), (C.c_double * cap)()
    k, s = (C.c_int32 * cap)(), (C.c_int32 * cap)()
    n = min(int(lib.cmt_timing_timeline(a, b, k, s, cap)), cap)
    return np.array(a[:n]), np.array(b[:n]), np.array(k[:n])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--molecules", default="1e7")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    lib = nat.lib()
    bl = lens_beamline(lens_table())
    src = eng.make_source(CeNTREXVelocityDistribution(), CeNTREXPositionDistribution())
    for size in args.molecules.split(","):
        n = int(float(size))
        prop = eng.Propagator(bl.elements, 0)
        ic = prop.draw(src, 2026, 0, n)
        for _ in range(3):
            prop.propagate_ic(ic, want_fate=True)
        torch.cuda.synchronize()
        lib.cmt_timing_enable(2)
        timeline(lib)
        for _ in range(args.reps):
            prop.propagate_ic(ic, want_fate=True)
        torch.cuda.synchronize()
        a, b, k = timeline(lib)
        lib.cmt_timing_enable(0)
        d = b - a
        out = {"molecules": n, "seg_ctas": os.environ.get("CMT_TUNE_SEG_CTAS", "default"),
               "walk_ms": float(np.median(d[k == 0])), "lens_stage_ms": float(np.median(d[k == 1])),
               "segments_ms": [float(np.median(d[k == 8 + j])) for j in range(16) if (k == 8 + j).any()],
               "tail_ms": float(np.median(d[k == 7]))}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
