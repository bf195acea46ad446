# round 2, call K: exact-mode step loop unrolled by two with alternating register sets (no moves at the end of a step)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2k_parity.log
for combo in 4:3 4:4 8:3 2:3; do
  c=${combo%%:*}; k=${combo##*:}
  CMT_TUNE_SEG_COPIES=$c CMT_TUNE_SEG_CTAS=$k timeout -s KILL 300 python profiles/ab_quick.py pingpong_copies${c}_ctas${k} 2>>gpurun_out/r2k.err | tee -a gpurun_out/r2k_ab.jsonl
done
