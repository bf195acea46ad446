# round 2, call L: GPU test suite, launch list, full ncu capture of one step (1e7) and of the lens segments at 8e7
T=r02l
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -15 > gpurun_out/${T}_tests.log; cat gpurun_out/${T}_tests.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-reference-python > gpurun_out/${T}_launches_bench.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'walk_kernel|lens_seg_kernel|tail_kernel' -s 12 -c 12 -f -o gpurun_out/${T}_full python profiles/prof_step.py > gpurun_out/${T}_full.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'lens_seg_kernel' -s 4 -c 4 -f -o gpurun_out/${T}_full_8e7 python profiles/prof_step.py 8e7 > gpurun_out/${T}_full_8e7.log 2>&1
ls -la gpurun_out/${T}_*
