#!/bin/bash
# build_variant.sh NAME [-DFLAG ...]: a variant of libcmt_b200.so into lib/variants/NAME.so (for A/B scripts)
set -e
cd "$(dirname "$0")/.."
L=centrex-molecule-trajectories_b200
NAME=$1; shift
mkdir -p $L/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -I include "$@" \
     -o $L/lib/variants/$NAME.so $L/csrc/cmt_api.cu
