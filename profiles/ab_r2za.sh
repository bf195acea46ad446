# round 2, call ZA: fewer ALU-pipe instructions in the round-2 step loop without touching its structure: the work
# counters of the reference path leave the common path (tail), the four range tests of the square roots go into
# the running-minimum record (range), both
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
for v in keep tail range both keep both; do
  if [ $v != keep ]; then cp $L/variants/$v.so $L/libcmt_b200.so; else cp /tmp/keep.so $L/libcmt_b200.so; fi
  timeout -s KILL 300 python profiles/ab_quick.py $v --slots 6 2>>gpurun_out/r2za.err | tee -a gpurun_out/r2za_ab.jsonl
done
cp /tmp/keep.so $L/libcmt_b200.so
