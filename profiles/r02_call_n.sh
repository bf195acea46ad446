# round 2, call N: both bench arms at N=1 on the current build
T=r02n
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2>gpurun_out/${T}_bench_reference.err
timeout -s KILL 900 python bench.py > gpurun_out/${T}_bench_n1.json 2>gpurun_out/${T}_bench_n1.err
tail -3 gpurun_out/${T}_bench_n1.err
python - <<P
import json
for f in ('gpurun_out/${T}_bench_reference.json','gpurun_out/${T}_bench_n1.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g' % d['value'], d.get('ms_per_step'), d.get('kernel_ms_per_step'), 'e2e %.4g' % d['e2e']['value'], {k: '%.4g' % d[k]['value'] for k in ('e2e_philox','e2e_api') if k in d}, (d.get('e2e_api') or {}).get('ms_per_call'), 'roof', (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('fp64_pipe_busy'), (d.get('roofline_walk') or {}).get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'launches', d.get('gpu_launches'), 'contracted', (d.get('contracted_math') or {}).get('value'), d.get('clocks'), d.get('host_binding'))
    except Exception as e: print(f, 'FAILED', e)
P
