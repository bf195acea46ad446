# round 2, call V: a lone Philox run cut into 2 / 3 / 4 pieces on as many streams; host-IC path with 2 / 4 staging slots
mkdir -p gpurun_out
for combo in 2:4 3:4 4:4 4:2; do
  pc=${combo%%:*}; sl=${combo##*:}
  CMT_TUNE_PHILOX_PIECES=$pc CMT_TUNE_IC_SLOTS=$sl timeout -s KILL 300 python profiles/ab_quick.py pieces${pc}_icslots${sl} --big 0 2>>gpurun_out/r2v.err | tee -a gpurun_out/r2v_ab.jsonl
done
CMT_TUNE_IC_SLOTS=2 timeout -s KILL 600 python bench.py --no-cpu --no-reference-python --no-contracted 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ic_slots 2: e2e %.4g' % d['e2e']['value'], 'value %.4g' % d['value'], 'philox %.4g' % d['e2e_philox']['value'])"
CMT_TUNE_IC_SLOTS=4 timeout -s KILL 600 python bench.py --no-cpu --no-reference-python --no-contracted 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ic_slots 4: e2e %.4g' % d['e2e']['value'], 'value %.4g' % d['value'], 'philox %.4g' % d['e2e_philox']['value'])"
