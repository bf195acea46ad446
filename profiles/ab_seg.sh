# A/B: lens integrator as a chain of segment launches (global re-packing of survivors between segments)
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --no-cpu --slots ${SLOTS:-3} ${AB_ARGS:-} > gpurun_out/abq_$name.json 2>gpurun_out/abq_$name.err
  env "$@" timeout -s KILL 300 python bench.py --no-cpu --molecules 8e7 --steps 5 --no-contracted --slots ${SLOTS:-3} > gpurun_out/abq8_$name.json 2>>gpurun_out/abq_$name.err
  python - <<P
import json
for f in ('gpurun_out/abq_$name.json','gpurun_out/abq8_$name.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); c=d.get('contracted_math') or {}
        print('$name', f[-24:], 'value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], d['kernel_ms_per_step'], 'philox %.4g' % d['e2e_philox']['value'], 'contracted', c.get('ms_per_step'), c.get('kernel_ms_per_step'))
    except Exception as e: print('$name', f, 'FAILED', e)
P
}
for s in 75 100 150 200 300 600; do
  run s${s}_c5 CMT_TUNE_SPLIT=1 CMT_TUNE_LENS_PRIO=1 CMT_TUNE_SEG=$s CMT_TUNE_SEG_CTAS=5
done
for c in 3 4; do
  run s150_c$c CMT_TUNE_SPLIT=1 CMT_TUNE_LENS_PRIO=1 CMT_TUNE_SEG=150 CMT_TUNE_SEG_CTAS=$c
done
SLOTS=4 run s150_c5_slots4 CMT_TUNE_SPLIT=1 CMT_TUNE_LENS_PRIO=1 CMT_TUNE_SEG=150 CMT_TUNE_SEG_CTAS=5
run s150_c5_noprio CMT_TUNE_SPLIT=1 CMT_TUNE_LENS_PRIO=0 CMT_TUNE_SEG=150 CMT_TUNE_SEG_CTAS=5
