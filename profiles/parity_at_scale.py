#!/usr/bin/env python
"""Parity of the CUDA path with the CPU oracle at sizes beyond the test suite.

    python profiles/parity_at_scale.py > profiles/r02_parity_at_scale.json
    python profiles/parity_at_scale.py --repeat 10 > profiles/r02_parity_at_scale_x10.json    # 10 seeds per case, one line each
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import oracle  # noqa: E402
from trajectories import _engine as eng  # noqa: E402
from trajectories.centrex import apertures_beamline, lens_beamline, lens_table, spa_beamline  # noqa: E402
from trajectories.distributions import (CeNTREXPositionDistribution, CeNTREXVelocityDistribution,  # noqa: E402
                                        GaussianPositionDistribution)


def compare(label, bl, vdist, xdist, n, seed, math="exact"):
    threads = oracle.host_cores()
    src_o = oracle.make_source(vdist, xdist)
    ic = oracle.draw(src_o, seed, 0, n, n_threads=threads)          # the oracle's own samples, shared by both sides
    t0 = time.perf_counter()
    want = oracle.propagate(bl.elements, ic, n_threads=threads)
    t_cpu = time.perf_counter() - t0
    prop = eng.Propagator(bl.elements, 0, math=math)
    prop.reset()
    res = prop.propagate_ic(torch.from_numpy(ic).cuda(), want_fate=True, want_final=True)
    torch.cuda.synchronize()
    fate = res.fate.cpu().numpy()
    fin = res.final.cpu().numpy()
    same = fate == want["fate"]
    rel = np.abs(fin[:, same] - want["fin"][:, same]) / np.maximum(np.abs(want["fin"][:, same]), 1e-9)
    bit = (fin[:, same].view(np.int64) == want["fin"][:, same].view(np.int64)).mean()
    # the same molecules again asking for fates only: that launch runs the walk kernel's FP32 fate filter
    prop.reset()
    res2 = prop.propagate_ic(torch.from_numpy(ic).cuda(), want_fate=True, want_final=False)
    torch.cuda.synchronize()
    fate2 = res2.fate.cpu().numpy()
    filt = dict(filter_decided=int(res2.work[5]), filter_fate_mismatches_vs_oracle=int((fate2 != want["fate"]).sum()),
                filter_fates_equal_unfiltered=bool((fate2 == fate).all()),
                filter_counters_equal=bool((res2.counters.cpu().numpy() == want["counters"]).all()),
                filter_rows_steps_equal=bool((res2.work[:2].cpu().numpy() == want["work"][:2]).all()))
    out = dict(case=label, math=math, seed=seed, molecules=n, **filt, lens_entries=int(want["work"][1] > 0) and int(res.work[3]),
               rk_steps=int(want["work"][1]), fate_mismatches=int((~same).sum()), max_rel_err=float(rel.max()),
               bit_identical_fraction=float(bit), counters_equal=bool((res.counters.cpu().numpy() == want["counters"]).all()),
               oracle_seconds=round(t_cpu, 2), oracle_threads=threads)
    print(json.dumps(out), flush=True)


def main():
    repeat = int(sys.argv[sys.argv.index("--repeat") + 1]) if "--repeat" in sys.argv else 1
    v, x = CeNTREXVelocityDistribution(), CeNTREXPositionDistribution()
    lens = lens_beamline(lens_table())
    for k in range(repeat):
        s = 100 * k             # seeds 11..14 for the first pass, 111..114 for the second, ...
        compare("lens beamline, standard source", lens, v, x, 50_000_000, 11 + s)
        compare("lens beamline, collimated source (most molecules enter the lens)", lens,
                CeNTREXVelocityDistribution(sigmax=3, sigmay=3), x, 4_000_000, 12 + s)
        compare("apertures-only beamline, standard source", apertures_beamline(), v, x, 50_000_000, 13 + s)
        compare("SPA beamline, Gaussian position source", spa_beamline(), v, GaussianPositionDistribution(), 50_000_000, 14 + s)
        if k == 0:
            compare("lens beamline, standard source", lens, v, x, 50_000_000, 11, math="contracted")
            compare("lens beamline, collimated source (most molecules enter the lens)", lens,
                    CeNTREXVelocityDistribution(sigmax=3, sigmay=3), x, 4_000_000, 12, math="contracted")


if __name__ == "__main__":
    main()
