# round 2, call ZG: the same RK step in other statement orders (CMT_ORDER bits: l2 before l1, l4 before l3, k3 before
# k2, the last two quotient records behind the position update) and under smaller register caps: what ptxas' list
# scheduler makes of the same operations
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
for v in keep o1 o2 o3 o4 o8 o15 r120 r112 keep; do
  if [ $v != keep ]; then cp $L/variants/$v.so $L/libcmt_b200.so; else cp /tmp/keep.so $L/libcmt_b200.so; fi
  timeout -s KILL 300 python profiles/ab_quick.py $v --slots 6 2>>gpurun_out/r2zg.err | tee -a gpurun_out/r2zg_ab.jsonl
done
cp /tmp/keep.so $L/libcmt_b200.so
