# A/B: lens grid sized for G molecules per lane (fewer, fuller warps that refill) x number of streams
mkdir -p gpurun_out
for g in 1 2 3 4; do for s in 3 4 6; do
  CMT_TUNE_WALK_CTAS=16 CMT_TUNE_LENS_PRIO=1 CMT_TUNE_LENS_GEN=$g python bench.py --no-cpu --no-contracted --slots $s > gpurun_out/abg_${g}_${s}.json 2>/dev/null
  python - <<P
import json
d=json.loads(open('gpurun_out/abg_${g}_${s}.json').read().strip().splitlines()[-1])
print('gen', $g, 'slots', $s, 'value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], d['kernel_ms_per_step'], 'philox %.4g' % d['e2e_philox']['value'])
P
done; done
