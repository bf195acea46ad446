#!/bin/bash
# Installs the UNMODIFIED reference (otimgren/centrex-molecule-trajectories) into baseline/_ref/ (git-ignored; it
# travels to the GPU box with the snapshot).  Run in the build container, where /root/reference exists:
#
#     bash baseline/install_ref.sh
#
#   baseline/_ref/trajectories/   the reference package, installed by pip from a scratch copy of the source tree
#                                 (the reference's setup.py writes egg-info into the tree, /root/reference is
#                                 read-only; --no-deps because h5py is neither installed nor in the wheelhouse)
#   baseline/_ref/examples/       the reference's example scripts, byte for byte (acceptance scripts: they must run
#                                 unchanged against this repository's package, tests/test_examples_dropin.py)
#
# Nothing under baseline/_ref is imported by the product.  bench.py's reference leg runs the installed package in
# a subprocess (baseline/run_reference.py) with oracle/stubs standing in for matplotlib, h5py, hexalattice and
# centrex_TlF, none of which exists in the image.
set -e
cd "$(dirname "$0")/.."
REF=${1:-/root/reference}
[ -d "$REF/src/trajectories" ] || { echo "reference tree not found at $REF"; exit 1; }
rm -rf /tmp/cmt_refcopy baseline/_ref
cp -r "$REF" /tmp/cmt_refcopy
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/cmt_refcopy
cp -r "$REF/examples" baseline/_ref/examples
rm -rf /tmp/cmt_refcopy
find baseline/_ref -name __pycache__ -type d -prune -exec rm -rf {} +
diff -r "$REF/src/trajectories" baseline/_ref/trajectories && diff -r "$REF/examples" baseline/_ref/examples && echo "baseline/_ref is identical to the reference sources"
