mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --no-cpu --no-contracted > gpurun_out/abt_$name.json 2>gpurun_out/abt_$name.err
  python - <<P
import json
for f in ('gpurun_out/abt_$name.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print('$name', 'value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], d['kernel_ms_per_step'], 'philox %.4g' % d['e2e_philox']['value'], 'api %.4g' % d['e2e_api']['value'], 'launches', d['gpu_launches'])
    except Exception as e: print('$name', f, 'FAILED', e)
P
}
for c in 3 4 5; do run spread_c$c CMT_TUNE_SEG_CTAS=$c; done
run spread_c3_s100 CMT_TUNE_SEG_CTAS=3 CMT_TUNE_SEG=100
run spread_c4_s100 CMT_TUNE_SEG_CTAS=4 CMT_TUNE_SEG=100
run spread_c4_s75 CMT_TUNE_SEG_CTAS=4 CMT_TUNE_SEG=75
