# round 2, call H: the GPU test suite on HEAD (no -x), then the quick A/B line
T=r02h
mkdir -p gpurun_out
timeout -s KILL 2400 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
timeout -s KILL 300 python profiles/ab_quick.py head 2>gpurun_out/${T}_ab.err | tee gpurun_out/${T}_ab.jsonl
