# round 2, call O (2 GPUs): the NCCL parity test and the bench line at N=2
T=r02o
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${T}_bench_n2.json 2>gpurun_out/${T}_bench_n2.err
tail -3 gpurun_out/${T}_bench_n2.err
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench_n2.json').read().strip().splitlines()[-1])
print('value %.4g' % d['value'], d.get('ms_per_step'), 'e2e %.4g' % d['e2e']['value'], {k: '%.4g' % d[k]['value'] for k in ('e2e_philox','e2e_api') if k in d}, d.get('host_binding'), d.get('clocks'))
P
