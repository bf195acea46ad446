"""The reference's own example scripts import and build their beamlines against this
package unchanged (same module paths, class names and keyword arguments).  They are
executed from /root/reference when that tree exists (build container only); without a
GPU they must stop at run_simulation with NativeError -- never fall back to a CPU path."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "centrex-molecule-trajectories_b200"
EXAMPLES = Path("/root/reference/examples")

pytestmark = pytest.mark.skipif(not EXAMPLES.exists(), reason="reference tree not present")


def run_example(script, *args):
    env = dict(os.environ, PYTHONPATH=f"{PKG}:{PKG / 'shims'}")
    return subprocess.run([sys.executable, str(script), *args], capture_output=True, text=True, env=env,
                          cwd=str(ROOT), timeout=300)


@pytest.mark.parametrize("script,args", [
    ("lens_simulation_beamline.py", ["--N_traj", "1e4"]),
    ("lens_simulation_different_states.py", []),
])
def test_example_reaches_the_gpu_call(script, args):
    import torch

    r = run_example(EXAMPLES / script, *args)
    if torch.cuda.is_available():
        pytest.skip("with a GPU the example runs on to the HDF save, which needs h5py")
    assert r.returncode != 0
    assert "NativeError" in r.stderr and "no CPU fallback" in r.stderr, r.stderr[-2000:]
    assert "run_simulation" in r.stderr                     # it got as far as the propagation call
    assert "ImportError" not in r.stderr and "TypeError" not in r.stderr
