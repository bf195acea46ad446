# round 2, call W: uneven halves of a lone Philox run (percent of the molecules in the first piece)
mkdir -p gpurun_out
for sp in 50 40 45 55 60 65 70; do
  CMT_TUNE_PHILOX_SPLIT=$sp timeout -s KILL 300 python profiles/ab_quick.py split_$sp --big 0 2>>gpurun_out/r2w.err | tee -a gpurun_out/r2w_ab.jsonl
done
