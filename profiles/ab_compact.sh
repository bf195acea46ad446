# A/B: 256-thread lens core kernel with in-CTA lane compaction every CHECK bursts; register caps 96/104/128
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
run() {
  name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --no-cpu --slots ${SLOTS:-3} > gpurun_out/abk_$name.json 2>gpurun_out/abk_$name.err
  env "$@" timeout -s KILL 300 python bench.py --no-cpu --molecules 8e7 --steps 5 --no-contracted --slots ${SLOTS:-3} > gpurun_out/abk8_$name.json 2>>gpurun_out/abk_$name.err
  python - <<P
import json
for f in ('gpurun_out/abk_$name.json','gpurun_out/abk8_$name.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d.get('contracted_math') or {}
        print('$name', f[-24:], 'value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], d['kernel_ms_per_step'], 'philox %.4g' % d['e2e_philox']['value'], 'contracted', c.get('ms_per_step'), c.get('kernel_ms_per_step'))
    except Exception as e: print('$name', f, 'FAILED', e)
P
}
for v in r96 r104 r128; do
  cp $L/variants/$v.so $L/libcmt_b200.so
  for k in 6 12 24 255; do
    run ${v}_k$k CMT_TUNE_SPLIT=1 CMT_TUNE_LENS_PRIO=1 CMT_TUNE_CHECK=$k
  done
done
cp $L/variants/r96.so $L/libcmt_b200.so
