# round 2, call X: configs[4] (1e10 molecules, one GPU, run_simulation): walk / lens CTAs per SM for 2^26-molecule launches
mkdir -p gpurun_out
for combo in 0:0 18:0 18:3 18:4 12:4 8:3 18:2; do
  w=${combo%%:*}; k=${combo##*:}
  echo "walk=$w seg=$k" | tee -a gpurun_out/r2x.log
  CMT_TUNE_WALK_CTAS=$w CMT_TUNE_SEG_CTAS=$k timeout -s KILL 300 python profiles/run_config5.py 2>>gpurun_out/r2x.err | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['math'], '%.4f s' % d['seconds'], '%.4g /s' % d['molecules_per_s'])
" | tee -a gpurun_out/r2x.log
done
