# A/B: CTAs per SM a lens segment launch asks for x streams of the overlapped pass (1e7 molecules per step)
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --no-cpu --no-contracted --slots ${SLOTS:-3} > gpurun_out/abr_$name.json 2>gpurun_out/abr_$name.err
  python - <<P
import json
for f in ('gpurun_out/abr_$name.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print('$name', 'value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], d['kernel_ms_per_step'], 'philox %.4g' % d['e2e_philox']['value'], 'api %.4g' % d['e2e_api']['value'], 'launches', d['gpu_launches'])
    except Exception as e: print('$name', f, 'FAILED', e)
P
}
for c in 2 3 4; do for sl in 3 4 6; do
  SLOTS=$sl run c${c}_slots$sl CMT_TUNE_SEG_CTAS=$c
done; done
SLOTS=4 run c3_slots4_w16 CMT_TUNE_SEG_CTAS=3 CMT_TUNE_WALK_CTAS=16
SLOTS=4 run c3_slots4_s120 CMT_TUNE_SEG_CTAS=3 CMT_TUNE_SEG=120
SLOTS=4 run c3_slots4_s200 CMT_TUNE_SEG_CTAS=3 CMT_TUNE_SEG=200
