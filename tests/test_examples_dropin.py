"""The reference's own example scripts run unchanged against this package (same module paths, class names and
keyword arguments).  The scripts are taken from /root/reference in the build container and from baseline/_ref/examples
(a byte-for-byte copy made by baseline/install_ref.sh, git-ignored, shipped to the GPU box) elsewhere.

Without a GPU they must stop at run_simulation with NativeError -- never fall back to a CPU path.  With a GPU they
run to their last line: simulate, print the Counter, save the result to HDF5 (through trajectories._minih5 when h5py is
not installed), and the file re-imports with the package's mirror of the reference's utils."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "centrex-molecule-trajectories_b200"
EXAMPLES = next((p for p in (Path("/root/reference/examples"), ROOT / "baseline" / "_ref" / "examples") if p.exists()), None)

pytestmark = pytest.mark.skipif(EXAMPLES is None, reason="reference examples not present (run baseline/install_ref.sh)")


def run_example(script, *args, cwd=ROOT):
    env = dict(os.environ, PYTHONPATH=f"{PKG}:{PKG / 'shims'}")
    return subprocess.run([sys.executable, str(script), *args], capture_output=True, text=True, env=env,
                          cwd=str(cwd), timeout=900)


@pytest.mark.parametrize("script,args", [
    ("lens_simulation_beamline.py", ["--N_traj", "1e4"]),
    ("lens_simulation_different_states.py", []),
])
def test_example_reaches_the_gpu_call(script, args):
    import torch

    if torch.cuda.is_available():
        pytest.skip("with a GPU the examples run to the end: test_example_runs_to_the_end")
    r = run_example(EXAMPLES / script, *args)
    assert r.returncode != 0
    assert "NativeError" in r.stderr and "no CPU fallback" in r.stderr, r.stderr[-2000:]
    assert "run_simulation" in r.stderr                     # it got as far as the propagation call
    assert "ImportError" not in r.stderr and "TypeError" not in r.stderr


@pytest.mark.gpu
def test_example_runs_to_the_end(tmp_path):
    """examples/lens_simulation_beamline.py, unchanged, default arguments (1e6 molecules, detected trajectories saved)."""
    (tmp_path / "saved_data").mkdir()                       # the script writes ./saved_data/lens_simulation_beamline.hdf
    r = run_example(EXAMPLES / "lens_simulation_beamline.py", cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-3000:]
    assert "Number of molecules that hit each element:" in r.stdout and "Beamline efficiency:" in r.stdout
    out = tmp_path / "saved_data" / "lens_simulation_beamline.hdf"
    assert out.read_bytes()[:8] == b"\x89HDF\r\n\x1a\n"
    for p in (str(ROOT), str(PKG)):
        if p not in sys.path:
            sys.path.insert(0, p)
    from trajectories import utils

    run = "Electrostatic lens simulation 4-7-2022 - det only - 1000000"
    res = utils.import_sim_result_from_hdf(out, run)
    assert sum(res.counter.counter_dict.values()) == 1_000_000
    assert len(res.molecules) == res.counter.counter_dict["Detected"] > 100
    assert [e.name for e in res.beamline.elements] == ["4K shield", "40K shield", "BB exit", "ES lens", "Field plates", "DR aperture"]
    m = res.molecules[0]
    assert m.aperture_hit == "Detected" and m.alive and m.trajectory.x.shape == (613, 3)
    assert np.isfinite(m.trajectory.x).all() and m.trajectory.x[-1, 2] > 6.44
