# round 2, call S: one lens CTA per SM per launch (each warp takes ~3 groups): lone stage and overlapped step, 4 / 6 / 8 streams
mkdir -p gpurun_out
for combo in 8:1:4 8:1:6 8:1:8 8:2:6 8:2:8; do
  w=${combo%%:*}; r=${combo#*:}; k=${r%%:*}; s=${r##*:}
  CMT_TUNE_WALK_CTAS=$w CMT_TUNE_SEG_CTAS=$k timeout -s KILL 300 python profiles/ab_quick.py walk${w}_seg${k}_slots${s} --slots $s --big 0 2>>gpurun_out/r2s.err | tee -a gpurun_out/r2s_ab.jsonl
done
