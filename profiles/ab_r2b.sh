# round 2, call B: FP64 operand microbenchmark + ncu --set full with source of the segment kernel at saturation
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_operands profiles/micro/fp64_operands.cu && /tmp/fp64_operands | tee gpurun_out/r2b_fp64_operands.txt
cp $L/libcmt_b200.so /tmp/keep.so
for v in new_c4; do
  cp $L/variants/$v.so $L/libcmt_b200.so
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'lens_seg_kernel' -s 4 -c 2 -f -o gpurun_out/r2b_${v}_8e7 python profiles/prof_step.py 8e7 > gpurun_out/r2b_${v}_8e7.log 2>&1
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'lens_seg_kernel' -s 8 -c 2 -f -o gpurun_out/r2b_${v}_1e7 python profiles/prof_step.py 1e7 > gpurun_out/r2b_${v}_1e7.log 2>&1
done
cp /tmp/keep.so $L/libcmt_b200.so
ls -la gpurun_out/r2b*
