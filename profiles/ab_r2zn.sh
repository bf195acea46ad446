# round 2, call ZN: segments of unequal length (most molecules that hit the bore do so early: short segments first re-pack
# the warps where lanes die, long ones later save launches)
mkdir -p gpurun_out
for plan in 150 75,75,150,300 50,100,150,300 100,100,100,300 75,75,150,150,150 75,75,450 100,200,300 50,50,100,100,300 150; do
  CMT_TUNE_SEG_PLAN=$plan timeout -s KILL 300 python profiles/ab_quick.py plan_$plan --slots 6 2>>gpurun_out/r2zn.err | tee -a gpurun_out/r2zn_ab.jsonl
done
