# A/B: co-residency of the walk kernel (FP32/HBM) and the lens kernel (FP64) in the overlapped pass.
mkdir -p gpurun_out
for w in 32 16 8 4; do for p in 0 1; do
  CMT_TUNE_WALK_CTAS=$w CMT_TUNE_LENS_PRIO=$p python bench.py --no-cpu --no-contracted ${AB_ARGS:-} > gpurun_out/abo_${w}_${p}.json 2>/dev/null
  python - <<P
import json
d=json.loads(open('gpurun_out/abo_${w}_${p}.json').read().strip().splitlines()[-1])
print('walk_ctas', $w, 'lens_prio', $p, 'value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], d['kernel_ms_per_step'], 'philox %.4g' % d['e2e_philox']['value'])
P
done; done
