# round 2, call A: parity of the new RK step, then quick A/B of lens-kernel variants (see profiles/ab_quick.py)
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 | tee gpurun_out/r2a_parity.log
for v in r01 new_u1 new_u2 new_c4 new_c6; do
  cp $L/variants/$v.so $L/libcmt_b200.so
  timeout -s KILL 300 python profiles/ab_quick.py $v 2>gpurun_out/r2a_$v.err | tee -a gpurun_out/r2a_ab.jsonl
done
cp /tmp/keep.so $L/libcmt_b200.so
