#!/usr/bin/env python
"""Quick A/B measurement of whatever library is in lib/libcmt_b200.so (one JSON line):

    python profiles/ab_quick.py [tag] [--molecules 1e7] [--steps 20] [--big 8e7]

  walk_ms, lens_ms     the walk kernel / the lens stage (segment launches + tail) alone: K steps back to back on
                       one stream, CUDA events around each kernel (cmt_timing_*), as bench.py's pass A
  step_ms              K steps on four alternating streams, one CUDA-graph replay per step (bench.py's pass B)
  philox_ms            the same through cmt_run_host_philox (what run_simulation does)
  big_lens_ms          the lens stage alone at --big molecules per launch (saturated regime)
  identical            fates, final rows and work counters of a lens-heavy sample equal those of the same library
                       forced onto the plain-intrinsic path (cmt_debug_flags(1)), bit for bit
"""
import argparse
import ctypes as C
import hashlib
import json
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


def timed_pass(lib, prop, ic, steps):
    lib.cmt_timing_enable(1)
    lib.cmt_timing_read(None, None, 1)
    prop.reset()
    for _ in range(steps):
        prop.propagate_ic(ic, want_fate=True)
    torch.cuda.synchronize()
    ms, nk = (C.c_double * 4)(), (C.c_int64 * 4)()
    lib.cmt_timing_read(ms, nk, 1)
    lib.cmt_timing_enable(0)
    return ms[0] / max(nk[0], 1), ms[1] / max(nk[1], 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag", nargs="?", default="lib")
    ap.add_argument("--molecules", type=float, default=1e7)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--big", type=float, default=8e7)
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--math", default="exact")
    args = ap.parse_args()
    lib = nat.lib()
    bl = lens_beamline(lens_table())
    src = eng.make_source(CeNTREXVelocityDistribution(), CeNTREXPositionDistribution())
    out = {"tag": args.tag, "math": args.math}

    # ---- bit identity against the plain-intrinsic path, lens-heavy sample ----
    rng = np.random.default_rng(77)
    n = 300_000
    ic = np.empty((6, n))
    th, rr = rng.uniform(0, 2 * np.pi, n), np.sqrt(rng.uniform(0, 1, n)) * 0.01
    ic[0], ic[1], ic[2] = rr * np.cos(th), rr * np.sin(th), 0.00635
    ic[3], ic[4], ic[5] = rng.normal(0, 3, n), rng.normal(0, 3, n), rng.normal(184, 16, n)
    ict = torch.from_numpy(ic).cuda()
    flat = eng.flatten(bl.elements)
    digests = []
    for flags in (0, 1):
        old = lib.cmt_debug_flags(flags)
        p = eng.Propagator(flat, 0, math=args.math)
        p.dev = eng.DeviceBeamline(flat, 0, args.math)
        p.reset()
        r = p.propagate_ic(ict, want_fate=True, want_final=True)
        torch.cuda.synchronize()
        h = hashlib.sha256()
        h.update(r.fate.cpu().numpy().tobytes())
        h.update(r.final.cpu().numpy().tobytes())
        w = r.work.cpu().numpy().copy()
        digests.append((h.hexdigest(), w[:4].tolist(), int(w[4])))
        lib.cmt_debug_flags(old)
    out["identical"] = digests[0][:2] == digests[1][:2] if args.math == "exact" else None
    out["digest"] = digests[0][0][:16]
    out["rk_steps"] = digests[0][1][1]
    out["rk_steps_on_reference_path"] = digests[0][2]

    # ---- timings ----
    nmol = int(args.molecules)
    prop = eng.Propagator(flat, 0, n_slots=args.slots, math=args.math)
    ic = prop.draw(src, 2026, 0, nmol)
    for _ in range(3):
        prop.reset()
        prop.propagate_ic(ic, want_fate=True)
    torch.cuda.synchronize()
    out["counters"] = prop.counters.cpu().tolist()
    out["walk_ms"], out["lens_ms"] = timed_pass(lib, prop, ic, args.steps)

    graphs = [prop.capture_ic(ic, want_fate=True, slot=s) for s in range(prop.n_slots)]
    for k in range(8):
        graphs[k % len(graphs)].replay()
    prop.join()
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.steps):
            graphs[k % len(graphs)].replay()
        prop.join()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        best = ms if best is None else min(best, ms)
    out["step_ms"] = best

    cnt = np.zeros(len(flat.fate_names), dtype=np.int64)
    work = np.zeros(8, dtype=np.int64)

    def philox():
        nat.check(lib.cmt_run_host_philox(prop.dev.handle, C.byref(src), 2026, 0, nmol, cnt.ctypes.data, work.ctypes.data))

    for _ in range(3):
        philox()
    import time
    t0 = time.perf_counter()
    for _ in range(args.steps):
        philox()
    out["philox_ms"] = 1e3 * (time.perf_counter() - t0) / args.steps
    del graphs, ic

    if args.big > 0:
        nbig = int(args.big)
        pb = eng.Propagator(flat, 0, math=args.math)
        icb = pb.draw(src, 2026, 0, nbig)
        for _ in range(2):
            pb.reset()
            pb.propagate_ic(icb, want_fate=True)
        torch.cuda.synchronize()
        out["big_walk_ms"], out["big_lens_ms"] = timed_pass(lib, pb, icb, 5)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
