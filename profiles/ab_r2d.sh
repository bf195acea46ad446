# round 2, call D: segment CTAs per SM x register budget x streams, new RK step
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
for v in new_u1 new_c4; do
  cp $L/variants/$v.so $L/libcmt_b200.so
  for c in 3 4 5; do
    for s in 4 6; do
      CMT_TUNE_SEG_CTAS=$c timeout -s KILL 300 python profiles/ab_quick.py ${v}_ctas${c}_slots${s} --slots $s 2>>gpurun_out/r2d.err | tee -a gpurun_out/r2d_ab.jsonl
    done
  done
done
cp /tmp/keep.so $L/libcmt_b200.so
