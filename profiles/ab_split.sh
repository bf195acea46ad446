# A/B: lens integrator split into a lean core kernel + tail kernel (CMT_TUNE_SPLIT), 4/5/6 core CTAs per SM
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
run() {  # name, env...
  name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --no-cpu ${AB_ARGS:-} > gpurun_out/abs_$name.json 2>gpurun_out/abs_$name.err
  env "$@" timeout -s KILL 300 python bench.py --no-cpu --molecules 8e7 --steps 5 --no-contracted > gpurun_out/abs8_$name.json 2>>gpurun_out/abs_$name.err
  python - <<P
import json
for f in ('gpurun_out/abs_$name.json','gpurun_out/abs8_$name.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d.get('contracted_math') or {}
        print('$name', f[-20:], 'value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], d['kernel_ms_per_step'], 'philox %.4g' % d['e2e_philox']['value'], 'contracted', c.get('ms_per_step'), c.get('kernel_ms_per_step'))
    except Exception as e: print('$name', f, 'FAILED', e)
P
}
cp $L/variants/core5.so $L/libcmt_b200.so
run old CMT_TUNE_SPLIT=0 CMT_TUNE_LENS_PRIO=1
for v in 4 5 6; do
  cp $L/variants/core$v.so $L/libcmt_b200.so
  run core$v CMT_TUNE_SPLIT=1 CMT_TUNE_LENS_PRIO=1
done
cp $L/variants/core5.so $L/libcmt_b200.so
