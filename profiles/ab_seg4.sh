# A/B: register budget of lens_seg_kernel (launch bound 3/4/5 CTAs per SM = up to 168/128/96 registers)
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
run() {
  name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --no-cpu > gpurun_out/abu_$name.json 2>gpurun_out/abu_$name.err
  env "$@" timeout -s KILL 300 python bench.py --no-cpu --molecules 8e7 --steps 5 --no-contracted --slots 3 > gpurun_out/abu8_$name.json 2>>gpurun_out/abu_$name.err
  python - <<P
import json
for f in ('gpurun_out/abu_$name.json','gpurun_out/abu8_$name.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d.get('contracted_math') or {}
        print('$name', f[-22:], 'value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], d['kernel_ms_per_step'], 'philox %.4g' % d['e2e_philox']['value'], 'api %.4g' % d['e2e_api']['value'], 'contracted', c.get('ms_per_step'), c.get('kernel_ms_per_step'))
    except Exception as e: print('$name', f, 'FAILED', e)
P
}
for m in 3 4 5; do
  cp $L/variants/m$m.so $L/libcmt_b200.so
  run m${m}_c3 CMT_TUNE_SEG_CTAS=3
done
cp $L/variants/m4.so $L/libcmt_b200.so
run m4_c4 CMT_TUNE_SEG_CTAS=4
cp $L/variants/m5.so $L/libcmt_b200.so
