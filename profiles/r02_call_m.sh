# round 2, call M: GPU suite on the hybrid / Hamiltonian / queue-capacity build, quick A/B line
T=r02m
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q --durations=6 2>&1 | tail -30 > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
timeout -s KILL 300 python profiles/ab_quick.py head_m 2>gpurun_out/${T}_ab.err | tee gpurun_out/${T}_ab.jsonl
