# round 2, call Q: walk CTAs per SM (grid = value x 148): few = leaves room for resident lens CTAs, many = short-lived CTAs
# that hand their slots to the higher-priority lens launches as they retire
mkdir -p gpurun_out
for w in 2 3 4 6 10 20 40 80 160 528; do
  CMT_TUNE_WALK_CTAS=$w timeout -s KILL 300 python profiles/ab_quick.py walkctas_$w --big 0 2>>gpurun_out/r2q.err | tee -a gpurun_out/r2q_ab.jsonl
done
