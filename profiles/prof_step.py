#!/usr/bin/env python
"""One bench step of configs[1] (1e7 molecules, lens beamline) in each source mode, for ncu:

    ncu --set full --clock-control none --import-source on -k regex:'walk_kernel|lens_seg_kernel|tail_kernel' \
        -s 12 -c 12 -o gpurun_out/prof python profiles/prof_step.py

launches: [warm-up] walk<ic>, 4 x lens_seg, tail, walk<philox>, 4 x lens_seg, tail, then the same twelve
again (captured)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "centrex-molecule-trajectories_b200")]

import torch  # noqa: E402

from trajectories import _engine as eng  # noqa: E402
from trajectories.centrex import lens_beamline, lens_table  # noqa: E402
from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
math = sys.argv[2] if len(sys.argv) > 2 else "exact"
bl = lens_beamline(lens_table())
prop = eng.Propagator(bl.elements, 0, math=math)
src = eng.make_source(CeNTREXVelocityDistribution(), CeNTREXPositionDistribution())
ic = prop.draw(src, 0, 0, n)
for _ in range(2):
    prop.propagate_ic(ic, want_fate=True)
    prop.propagate_philox(src, 0, 0, n)
torch.cuda.synchronize()
print(dict(zip(prop.flat.fate_names, prop.counters.cpu().tolist())), prop.work.cpu().tolist())
