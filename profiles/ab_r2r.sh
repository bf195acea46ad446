# round 2, call R: lone Philox run (cmt_run_host_philox: two halves on two streams): walk CTAs per SM x lens CTAs per SM
mkdir -p gpurun_out
for combo in 5:3 6:3 7:3 8:3 6:2 8:2 18:2 6:4; do
  w=${combo%%:*}; k=${combo##*:}
  CMT_TUNE_WALK_CTAS=$w CMT_TUNE_SEG_CTAS=$k timeout -s KILL 300 python profiles/ab_quick.py walk${w}_seg${k} --big 0 2>>gpurun_out/r2r.err | tee -a gpurun_out/r2r_ab.jsonl
done
