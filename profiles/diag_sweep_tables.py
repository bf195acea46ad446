"""How long do the 40 Stark tables of configs[2] take on this box, and how does that scale with threads?"""
import os, sys, time
from copy import copy
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")]
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from threadpoolctl import threadpool_limits
from trajectories.centrex import lens_beamline
from trajectories import stark_potential as sp, _tlf_full as T

bl = lens_beamline(); lens = bl.find_element("ES lens")
states = [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1)]
Vs = [20e3, 24e3, 27.6e3, 30e3, 34e3]
pts = [(s, V) for s in states for V in Vs]
def table_of(point):
    probe = copy(lens); probe.state, probe.V, probe.a_interp = point[0], point[1], None
    return probe.ensure_a_interp()
table_of(pts[0])
print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)))
for limit in (1, None):
    for w in (1, 2, 4, 8, 16):
        sp._CURVES.clear(); T._FOLLOWED.clear(); T._FOLLOWED_LOCKS.clear()
        t = time.perf_counter()
        ctx = threadpool_limits(limit) if limit else threadpool_limits(None)
        with ctx, ThreadPoolExecutor(max_workers=w) as pool:
            list(pool.map(table_of, pts))
        print("blas_limit", limit, "workers", w, "%.3f s" % (time.perf_counter() - t), flush=True)
# one block alone
C = T.XConstants(); E = np.linspace(0, 30000, 222)
for mf in (0, 2, 4):
    T._FOLLOWED.clear(); t = time.perf_counter(); T._followed_block(mf, E, 1e-4, 100.0, 6, C); print("block", mf, "%.4f s" % (time.perf_counter() - t))
idx, H0, HS, HZ = T._block(0, 6, C)
A = (H0 + 1e-4 * HZ)[None] + E[:, None, None] * HS[None]
t = time.perf_counter(); np.linalg.eigh(A); print("eigh 222x26x26 %.4f s" % (time.perf_counter() - t))
import scipy.linalg as sl
t = time.perf_counter()
for a in A: sl.eigh(a, driver="evr", check_finite=False)
print("scipy evr loop %.4f s" % (time.perf_counter() - t))
t = time.perf_counter()
for a in A: sl.eigh(a, driver="evd", check_finite=False)
print("scipy evd loop %.4f s" % (time.perf_counter() - t))
