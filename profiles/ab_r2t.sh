# round 2, call T: segment length re-tuned on the replicated-table kernel with 2 CTAs per SM per launch
mkdir -p gpurun_out
for g in 75 100 120 150 200 300; do
  CMT_TUNE_SEG=$g timeout -s KILL 300 python profiles/ab_quick.py seg_$g --slots 6 --big 0 2>>gpurun_out/r2t.err | tee -a gpurun_out/r2t_ab.jsonl
done
