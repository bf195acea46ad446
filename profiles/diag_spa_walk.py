"""Walk-kernel time of the SPA beamline (configs[3]: rectangular apertures, field plates, Gaussian position source)."""
import ctypes as C, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")]
import torch
from trajectories import _engine as eng, _native as nat
from trajectories.centrex import spa_beamline, lens_beamline, lens_table
from trajectories.distributions import GaussianPositionDistribution, CeNTREXPositionDistribution, CeNTREXVelocityDistribution
lib = nat.lib()
for name, bl, xd in (("spa/gauss", spa_beamline(), GaussianPositionDistribution()), ("spa/disc", spa_beamline(), CeNTREXPositionDistribution()),
                     ("lens/disc", lens_beamline(lens_table()), CeNTREXPositionDistribution()), ("lens/gauss", lens_beamline(lens_table()), GaussianPositionDistribution())):
    src = eng.make_source(CeNTREXVelocityDistribution(), xd)
    prop = eng.Propagator(bl.elements, 0)
    for n in (10_000_000, 1 << 26):
        prop.propagate_philox(src, 7, 0, n); torch.cuda.synchronize()
        prop.reset()
        lib.cmt_timing_enable(1); lib.cmt_timing_read(None, None, 1)
        t = time.perf_counter()
        for r in range(3):
            prop.propagate_philox(src, 7, r * n, n)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t) / 3
        ms, nk = (C.c_double * 4)(), (C.c_int64 * 4)()
        lib.cmt_timing_read(ms, nk, 1); lib.cmt_timing_enable(0)
        w = prop.work.cpu().numpy()
        print(name, n, "walk %.3f ms" % (ms[0] / max(nk[0], 1)), "lens %.3f ms" % (ms[1] / max(nk[1], 1)), "wall %.3f ms" % (1e3 * wall),
              "filtered %.4f" % (w[5] / (3 * n)), "counters", prop.counters.cpu().numpy().tolist(), flush=True)
