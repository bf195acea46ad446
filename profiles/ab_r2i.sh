# round 2, call I: the replicated shared-memory table of lens_seg_kernel: copies 1 / 2 / 4 / 8, 3 or 4 CTAs per SM per launch
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2i_parity.log
for combo in 1:3 2:3 4:3 4:4 8:3 8:2; do
  c=${combo%%:*}; k=${combo##*:}
  CMT_TUNE_SEG_COPIES=$c CMT_TUNE_SEG_CTAS=$k timeout -s KILL 300 python profiles/ab_quick.py copies${c}_ctas${k} 2>>gpurun_out/r2i.err | tee -a gpurun_out/r2i_ab.jsonl
done
