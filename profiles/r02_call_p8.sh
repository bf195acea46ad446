# round 2 (8 GPUs): the bench line under torchrun at N=8, and the reference arm launched the same way
T=r02p8
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/${T}_bench_n8.json 2>gpurun_out/${T}_bench_n8.err
tail -3 gpurun_out/${T}_bench_n8.err
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_n8.json 2>gpurun_out/${T}_bench_reference_n8.err
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench_n8.json').read().strip().splitlines()[-1])
print('value %.4g' % d['value'], d.get('ms_per_step'), 'e2e %.4g' % d['e2e']['value'], {k: '%.4g' % d[k]['value'] for k in ('e2e_philox','e2e_api') if k in d}, d.get('host_binding'), d.get('clocks'))
r=[l for l in open('gpurun_out/${T}_bench_reference_n8.json').read().strip().splitlines() if l.startswith('{')]
print(len(r), 'reference line(s)', r[-1][:200])
P
nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i "numa\|socket\|^CPU(s)" | head
