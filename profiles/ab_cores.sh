# A/B: co-residency of walk CTAs with the 96-register lens core kernel
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --no-cpu --no-contracted --slots ${SLOTS:-3} > gpurun_out/abc_$name.json 2>gpurun_out/abc_$name.err
  env "$@" timeout -s KILL 300 python bench.py --no-cpu --molecules 8e7 --steps 5 --no-contracted --slots ${SLOTS:-3} > gpurun_out/abc8_$name.json 2>>gpurun_out/abc_$name.err
  python - <<P
import json
for f in ('gpurun_out/abc_$name.json','gpurun_out/abc8_$name.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print('$name', f[-22:], 'value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], d['kernel_ms_per_step'], 'philox %.4g' % d['e2e_philox']['value'])
    except Exception as e: print('$name', f, 'FAILED', e)
P
}
for c in 3 4 5; do for w in 32 8 4; do for wp in 0 1; do
  run c${c}_w${w}_p${wp} CMT_TUNE_SPLIT=1 CMT_TUNE_LENS_PRIO=$((1-wp)) CMT_TUNE_WALK_PRIO=$wp CMT_TUNE_CORE_CTAS=$c CMT_TUNE_WALK_CTAS=$w
done; done; done
