# Round-end evidence run (one gpurun call): GPU tests, smoke, both bench arms, ncu launch list + full captures, timeline, sanitizers.
# usage: bash profiles/round_end.sh <tag>      (outputs land in gpurun_out/<tag>_*)
T=${1:-r02z}
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -12 > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout -s KILL 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2>gpurun_out/${T}_bench_reference.err
timeout -s KILL 900 python bench.py > gpurun_out/${T}_bench_n1.json 2>gpurun_out/${T}_bench_n1.err
python - <<P
import json
for f in ('gpurun_out/${T}_bench_reference.json','gpurun_out/${T}_bench_n1.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value %.4g' % d['value'], d.get('ms_per_step'), d.get('kernel_ms_per_step'), 'e2e %.4g' % d['e2e']['value'], {k: '%.4g' % d[k]['value'] for k in ('e2e_philox','e2e_api') if k in d}, 'roof', (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('fp64_pipe_busy'), (d.get('roofline_one_stream') or {}).get('frac'), (d.get('roofline_walk') or {}).get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'launches', d.get('gpu_launches'), 'contracted', (d.get('contracted_math') or {}).get('value'))
    except Exception as e: print(f, 'FAILED', e)
P
timeout -s KILL 300 python profiles/timeline_pass_b.py --out gpurun_out/${T}_timeline_pass_b.json > /dev/null 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-reference-python > gpurun_out/${T}_launches_bench.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'walk_kernel|lens_seg_kernel|tail_kernel' -s 12 -c 12 -f -o gpurun_out/${T}_full python profiles/prof_step.py > gpurun_out/${T}_full.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none -k regex:'lens_seg_kernel' -s 4 -c 4 -f -o gpurun_out/${T}_full_8e7 python profiles/prof_step.py 8e7 > gpurun_out/${T}_full_8e7.log 2>&1
timeout -s KILL 600 ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:'walk_kernel|lens_seg_kernel|tail_kernel' -s 12 -c 12 --csv --log-file gpurun_out/${T}_fp64_inst.csv python profiles/prof_step.py > /dev/null 2>&1
timeout -s KILL 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_hybrid.py tests/test_gpu_queue.py -q -x -k "golden or unusual or hostile or early or resume or small_queue or first_last" > gpurun_out/${T}_memcheck.log 2>&1; tail -4 gpurun_out/${T}_memcheck.log
timeout -s KILL 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_racecheck.log 2>&1; tail -3 gpurun_out/${T}_racecheck.log
ls -la gpurun_out/${T}_*
