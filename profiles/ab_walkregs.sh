# A/B: register budget of walk_kernel (launch bound 16/24/32 CTAs of 64 threads per SM = 62/40/32 registers)
mkdir -p gpurun_out
L=centrex-molecule-trajectories_b200/lib
cp $L/libcmt_b200.so /tmp/keep.so
run() {
  name=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --no-cpu --no-contracted > gpurun_out/abw_$name.json 2>gpurun_out/abw_$name.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/abw_$name.json').read().strip().splitlines()[-1])
    print('$name', 'value %.4g' % d['value'], 'ms %.4f' % d['ms_per_step'], d['kernel_ms_per_step'], 'philox %.4g' % d['e2e_philox']['value'], 'walkfrac %.3f' % d['roofline_walk']['frac'])
except Exception as e: print('$name', 'FAILED', e)
P
}
for v in 16 24 32; do
  cp $L/variants/w$v.so $L/libcmt_b200.so
  for g in 16 24 32; do
    [ $g -le $v ] || [ $v -eq 16 ] && run w${v}_g$g CMT_TUNE_WALK_CTAS=$g
  done
done
cp /tmp/keep.so $L/libcmt_b200.so
