"""Live cross-check in the build container: the unmodified reference (imported from
/root/reference through oracle/stubs, in a subprocess so its `trajectories` package does not
clash with ours) against the C oracle on fresh random initial conditions.  Skipped where the
reference tree is absent (e.g. on the GPU box); the committed fixtures cover that case."""
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/src/trajectories")

pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference tree not present")

SCRIPT = textwrap.dedent("""
    import sys
    import numpy as np
    sys.path.insert(0, "{root}/tests/golden")
    import make_golden as mg                      # sets up the stub packages + reference import
    rng = np.random.default_rng({seed})
    n = {n}
    ic = np.empty((6, n))
    th, rr = rng.uniform(0, 2 * np.pi, n), np.sqrt(rng.uniform(0, 1, n)) * 0.01
    ic[0], ic[1], ic[2] = rr * np.cos(th), rr * np.sin(th), 0.00635
    ic[3], ic[4], ic[5] = rng.normal(0, {sigma}, n), rng.normal(0, {sigma}, n), rng.normal(184, 16, n)
    table = mg.lens_table(J={J}, mJ={mJ}, V={V})
    res = mg.run_reference(mg.lens_beamline(table), ic)
    np.savez("{out}", ic=ic, table_r=table[0], table_a=table[1], **res)
""")


@pytest.mark.parametrize("seed,n,sigma,J,mJ,V", [(101, 400, 39.5, 2, 0, 27.6e3), (102, 60, 3.0, 3, 1, 32e3)])
def test_reference_equals_oracle_bit_for_bit(tmp_path, seed, n, sigma, J, mJ, V):
    from oracle import oracle
    from tests.beamlines import lens_beamline

    out = tmp_path / "live.npz"
    code = SCRIPT.format(root=ROOT, seed=seed, n=n, sigma=sigma, J=J, mJ=mJ, V=V, out=out)
    subprocess.run([sys.executable, "-c", code], check=True, cwd=str(ROOT), timeout=600)
    g = np.load(out)
    res = oracle.propagate(lens_beamline((g["table_r"], g["table_a"])).elements, g["ic"])
    assert res["fate_names"] == list(g["fate_names"])
    np.testing.assert_array_equal(res["fate"], g["fate"])
    np.testing.assert_array_equal(res["n_rows"], g["n_rows"])
    np.testing.assert_array_equal(res["fin"].view(np.int64), g["fin"].view(np.int64))
