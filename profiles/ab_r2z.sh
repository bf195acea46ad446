# round 2, call Z: RK steps in blocks with a sticky validity record (in-place update, no per-step commit) and the
# table entry by masking; GPU tests first, then A/B of the variants built by build_variant.sh
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2z_tests.log
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
for v in new old blk10_nomask blk0_mask blk5 blk25 blk150 new; do
  if [ $v != new ]; then cp $L/variants/$v.so $L/libcmt_b200.so; else cp /tmp/keep.so $L/libcmt_b200.so; fi
  timeout -s KILL 300 python profiles/ab_quick.py $v --slots 6 2>>gpurun_out/r2z.err | tee -a gpurun_out/r2z_ab.jsonl
done
cp /tmp/keep.so $L/libcmt_b200.so
tail -3 gpurun_out/r2z_tests.log
