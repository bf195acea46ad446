# round 2, call J: with the table's bank conflicts gone, do more resident warps pay?  96 / 80 registers (5 / 6 CTAs per SM)
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
for combo in m5:4:3 m5:4:4 m5:4:5 m5:2:5 m6:4:4 m6:4:6 m6:2:6; do
  v=${combo%%:*}; r=${combo#*:}; c=${r%%:*}; k=${r##*:}
  cp $L/variants/$v.so $L/libcmt_b200.so
  CMT_TUNE_SEG_COPIES=$c CMT_TUNE_SEG_CTAS=$k timeout -s KILL 300 python profiles/ab_quick.py ${v}_copies${c}_ctas${k} 2>>gpurun_out/r2j.err | tee -a gpurun_out/r2j_ab.jsonl
done
cp /tmp/keep.so $L/libcmt_b200.so
