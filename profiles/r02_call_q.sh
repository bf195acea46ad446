# round 2, call Q: after the shared slot streams and the shared Stark eigenpairs: GPU tests, bench (both arms), the BASELINE
# configurations through the public API, the API profile
T=r02q
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
timeout -s KILL 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2>gpurun_out/${T}_bench_reference.err
timeout -s KILL 900 python bench.py > gpurun_out/${T}_bench_n1.json 2>gpurun_out/${T}_bench_n1.err
timeout -s KILL 900 python profiles/run_configs.py > gpurun_out/${T}_configs.jsonl 2>gpurun_out/${T}_configs.err
timeout -s KILL 300 python profiles/prof_api.py > gpurun_out/${T}_prof_api.txt 2>&1
python - <<P
import json
for f in ('gpurun_out/${T}_bench_reference.json','gpurun_out/${T}_bench_n1.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g' % d['value'], d.get('ms_per_step'), 'e2e %.4g' % d['e2e']['value'], {k: '%.4g' % d[k]['value'] for k in ('e2e_philox','e2e_api') if k in d}, 'roof', (d.get('roofline') or {}).get('frac'), (d.get('roofline_one_stream') or {}).get('frac'), (d.get('roofline_walk') or {}).get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'FAILED', e)
for l in open('gpurun_out/${T}_configs.jsonl'):
    d=json.loads(l); print(d['config'][:60], '%.4f s' % d['seconds'], '%.3g /s' % d['molecules_per_s'])
P
grep "no saving\|^call" gpurun_out/${T}_prof_api.txt
