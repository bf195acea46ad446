# round 2, call P: how does the overlapped step scale with the launch size (is the 1e7 step short of the saturated lens stage
# because of small launches, or because of the walk kernel beside it?), and with the number of streams
mkdir -p gpurun_out
for m in 1e7 2e7 4e7 8e7; do
  timeout -s KILL 300 python profiles/ab_quick.py size_$m --molecules $m --steps 8 --big 0 2>>gpurun_out/r2p.err | tee -a gpurun_out/r2p_ab.jsonl
done
for s in 2 3 6 8; do
  timeout -s KILL 300 python profiles/ab_quick.py slots_$s --slots $s --big 0 2>>gpurun_out/r2p.err | tee -a gpurun_out/r2p_ab.jsonl
done
