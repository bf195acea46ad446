# round 2, call U: lens CTA size: 64 threads (8 per SM fit) and 256 threads (2 per SM) against 128
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
for combo in t64:2:4 t64:2:6 t64:4:4 t256:8:1 t256:4:1 t256:4:2; do
  v=${combo%%:*}; r=${combo#*:}; c=${r%%:*}; k=${r##*:}
  cp $L/variants/$v.so $L/libcmt_b200.so
  CMT_TUNE_SEG_COPIES=$c CMT_TUNE_SEG_CTAS=$k timeout -s KILL 300 python profiles/ab_quick.py ${v}_copies${c}_ctas${k} --slots 6 2>>gpurun_out/r2u.err | tee -a gpurun_out/r2u_ab.jsonl
done
cp /tmp/keep.so $L/libcmt_b200.so
