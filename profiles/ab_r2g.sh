# round 2, call G: lean step loop (z, t, vz, index out of the loop) at 116 / 96 / 80 / 72 registers x CTAs per SM per launch
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
for combo in lean4:3 lean4:4 lean5:3 lean5:5 lean6:3 lean6:4 lean6:6 lean7:3 lean7:5 lean7:7; do
  v=${combo%%:*}; c=${combo##*:}
  cp $L/variants/$v.so $L/libcmt_b200.so
  CMT_TUNE_SEG_CTAS=$c timeout -s KILL 300 python profiles/ab_quick.py ${v}_ctas${c} 2>>gpurun_out/r2g.err | tee -a gpurun_out/r2g_ab.jsonl
done
cp /tmp/keep.so $L/libcmt_b200.so
