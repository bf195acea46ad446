#!/usr/bin/env python
"""Timeline of the bench's overlapped pass (consecutive 1e7-molecule steps on alternating streams), kernel by kernel,
from CUDA events on the launch streams (cmt_timing_enable(2) / cmt_timing_timeline) -- the stand-in for a system
profiler (nsys is not in the image; ncu serialises kernels, so it cannot show this regime).

    python profiles/timeline_pass_b.py [--molecules 1e7] [--steps 24] [--slots 6] [--out profiles/r02_timeline_pass_b.json]

Prints one JSON object: per kernel kind the number of launches, the mean / min / max duration while overlapped and
the duration of the same kernel alone (one stream); the span of the pass; the time-weighted number of lens-segment
kernels and walk kernels in flight; how much of the span has no lens kernel in flight; and the first steps' intervals
for plotting.  The event pairs cost ~2 us per kernel (plain launches, no CUDA graph), so the step here is a few per
cent slower than the bench's graph replays; the shares are what the picture is for."""
import argparse
import ctypes as C
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

KINDS = {0: "walk_kernel", 1: "lens stage (4 segments + tail)", 7: "tail_kernel", 8: "lens_seg_kernel 0",
         9: "lens_seg_kernel 1", 10: "lens_seg_kernel 2", 11: "lens_seg_kernel 3"}


def timeline(lib, cap=1 << 16):
    a, b = (C.c_double * cap)(), (C.c_double * cap)()
    k, s = (C.c_int32 * cap)(), (C.c_int32 * cap)()
    n = lib.cmt_timing_timeline(a, b, k, s, cap)
    n = min(int(n), cap)
    return np.array(a[:n]), np.array(b[:n]), np.array(k[:n]), np.array(s[:n])


def in_flight(a, b, t_lo, t_hi):
    """time-weighted mean number of the intervals [a, b] that are open inside [t_lo, t_hi], and the idle share"""
    pts = sorted([(x, 1) for x in a] + [(x, -1) for x in b])
    level, last, area, idle = 0, t_lo, 0.0, 0.0
    for t, d in pts:
        t = min(max(t, t_lo), t_hi)
        area += level * (t - last)
        if level == 0:
            idle += t - last
        last = t
        level += d
    idle += max(0.0, t_hi - last) if level == 0 else 0.0
    span = max(t_hi - t_lo, 1e-12)
    return area / span, idle / span


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--molecules", type=float, default=1e7)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--slots", type=int, default=6)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    lib = nat.lib()
    n = int(args.molecules)
    bl = lens_beamline(lens_table())
    src = eng.make_source(CeNTREXVelocityDistribution(), CeNTREXPositionDistribution())
    prop = eng.Propagator(bl.elements, 0, n_slots=args.slots)
    ic = prop.draw(src, 2026, 0, n)
    for k in range(2 * args.slots):
        prop.propagate_ic(ic, want_fate=True, slot=k)
    prop.join()
    torch.cuda.synchronize()

    out = {"molecules_per_step": n, "steps": args.steps, "streams": args.slots,
           "what": "CUDA-event timeline of consecutive steps on alternating streams (plain launches)"}
    # alone: one stream
    lib.cmt_timing_enable(2)
    timeline(lib)
    for _ in range(4):
        prop.propagate_ic(ic, want_fate=True)
    torch.cuda.synchronize()
    a, b, k, s = timeline(lib)
    alone = {KINDS[q]: float(np.mean((b - a)[k == q])) for q in KINDS if (k == q).any()}
    # overlapped
    prop.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(args.steps):
        prop.propagate_ic(ic, want_fate=True, slot=j)
    prop.join()
    e1.record()
    torch.cuda.synchronize()
    a, b, k, s = timeline(lib)
    lib.cmt_timing_enable(0)
    out["ms_per_step_with_event_pairs"] = e0.elapsed_time(e1) / args.steps
    # steady state: drop the first and last `slots` steps (pipeline fill and drain)
    order = np.argsort(a[k == 0])
    walk_starts = a[k == 0][order]
    t_lo = float(walk_starts[min(args.slots, len(walk_starts) - 1)])
    t_hi = float(np.sort(b[k == 7])[max(len(walk_starts) - args.slots - 1, 0)])
    out["window_ms"] = [t_lo, t_hi]
    steps_in_window = int(((a[k == 0] >= t_lo) & (a[k == 0] < t_hi)).sum())
    out["steps_started_in_window"] = steps_in_window
    out["ms_per_step_in_window"] = (t_hi - t_lo) / max(steps_in_window, 1)
    kinds = {}
    for q, name in KINDS.items():
        m = (k == q) & (a >= t_lo) & (b <= t_hi)
        if m.any():
            d = (b - a)[m]
            kinds[name] = {"launches": int(m.sum()), "mean_ms": float(d.mean()), "min_ms": float(d.min()), "max_ms": float(d.max()),
                           "alone_ms": alone.get(name)}
    out["kernels"] = kinds
    seg = (k >= 8)
    lens_level, lens_idle = in_flight(a[seg], b[seg], t_lo, t_hi)
    walk_level, walk_idle = in_flight(a[k == 0], b[k == 0], t_lo, t_hi)
    out["lens_segment_kernels_in_flight_mean"] = lens_level
    out["share_of_time_without_a_lens_segment_kernel"] = lens_idle
    out["walk_kernels_in_flight_mean"] = walk_level
    out["share_of_time_without_a_walk_kernel"] = walk_idle
    # gaps between consecutive kernels of one stream (launch latency + event pair)
    gaps = []
    for sid in np.unique(s):
        m = (s == sid) & (k != 1)
        aa, bb = a[m], b[m]
        o = np.argsort(aa)
        gaps += list(aa[o][1:] - bb[o][:-1])
    gaps = np.array(gaps)
    out["gap_between_consecutive_kernels_of_a_stream_ms"] = {"median": float(np.median(gaps)), "p90": float(np.percentile(gaps, 90))}
    first = np.argsort(a)[: 7 * 2 * args.slots]
    out["first_intervals"] = [{"kind": KINDS.get(int(k[i]), str(int(k[i]))), "stream": int(s[i]), "start_ms": round(float(a[i]), 4),
                               "end_ms": round(float(b[i]), 4)} for i in first if k[i] != 1]
    text = json.dumps(out)
    print(text)
    if args.out:
        Path(args.out).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
