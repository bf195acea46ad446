mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so $L/variants/default.so
for v in default idx64 cidx32; do
  cp $L/variants/$v.so $L/libcmt_b200.so
  python bench.py --no-cpu > gpurun_out/ab_$v.json 2>/dev/null
  python bench.py --no-cpu --molecules 8e7 --steps 5 > gpurun_out/ab8_$v.json 2>/dev/null
  python -c "
import json
for f in ('gpurun_out/ab_$v.json','gpurun_out/ab8_$v.json'):
    d=json.load(open(f)); c=d['contracted_math']
    print('$v', f[-14:], 'exact', round(d['ms_per_step'],4), d['kernel_ms_per_step'], 'contracted', round(c['ms_per_step'],4), c['kernel_ms_per_step'], d['work_per_step']['rk_steps_on_reference_path'])
"
done
cp $L/variants/default.so $L/libcmt_b200.so
python -m pytest tests/test_gpu_parity.py -x -q -k "fast_math or contracted" 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"lens_kernel" -s 2 -c 1 -f -o gpurun_out/prof_r1g python profiles/prof_step.py > gpurun_out/prof_step_g.log 2>&1
