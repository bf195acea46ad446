# round 2, call ZL: the block-wise step loop again, now with the parked state in shared memory and the per-step path as
# its fallback, and the two force evaluations of a half written side by side (tight pairing in ptxas' schedule)
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
for v in keep pair pair_blk10 pair_blk5 pair_blk25 blk10 keep pair_blk10; do
  if [ $v != keep ]; then cp $L/variants/$v.so $L/libcmt_b200.so; else cp /tmp/keep.so $L/libcmt_b200.so; fi
  timeout -s KILL 300 python profiles/ab_quick.py $v --slots 6 2>>gpurun_out/r2zl.err | tee -a gpurun_out/r2zl_ab.jsonl
done
cp /tmp/keep.so $L/libcmt_b200.so
