# round 2, call ZM: why is a variant with the same FP64 instructions slower?  ncu --set full of the first lens segment at 8e7
# molecules with 4 CTAs per SM (saturated), committed library against the paired-evaluation variant
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
for v in keep pair; do
  if [ $v != keep ]; then cp $L/variants/$v.so $L/libcmt_b200.so; else cp /tmp/keep.so $L/libcmt_b200.so; fi
  CMT_TUNE_SEG_CTAS=4 timeout -s KILL 600 ncu --set full --clock-control none -k regex:'lens_seg_kernel' -s 4 -c 1 -f -o gpurun_out/r2zm_$v python profiles/prof_step.py 8e7 > gpurun_out/r2zm_$v.log 2>&1
done
cp /tmp/keep.so $L/libcmt_b200.so
ls -la gpurun_out/r2zm_*
