# round 2, call Y: what does an integer instruction cost next to the FP64 stream of the RK step?  4 / 8 dummy rounds of
# (SHF, LOP3, IADD) per step = +12 / +24 instructions on 241
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
for v in keep dummy4 dummy8; do
  if [ $v != keep ]; then cp $L/variants/$v.so $L/libcmt_b200.so; fi
  timeout -s KILL 300 python profiles/ab_quick.py $v --slots 6 2>>gpurun_out/r2y.err | tee -a gpurun_out/r2y_ab.jsonl
  CMT_TUNE_SEG_CTAS=4 timeout -s KILL 300 python profiles/ab_quick.py ${v}_seg4 --slots 6 2>>gpurun_out/r2y.err | tee -a gpurun_out/r2y_ab.jsonl
done
cp /tmp/keep.so $L/libcmt_b200.so
