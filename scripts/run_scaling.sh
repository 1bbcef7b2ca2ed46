#!/bin/bash
# 1 -> N GPU scaling run of bench.py plus configs 4/5 of BASELINE.json (dev tool; run under gpurun --gpus N)
# usage: run_scaling.sh [N=8] [list of GPU counts="1 2 4 8"]
N=${1:-8}
LIST=${2:-"1 2 4 8"}
mkdir -p gpurun_out
for n in $LIST; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
  tail -c 300 gpurun_out/scale_n$n.err | grep -v OMP_NUM | tail -2
done
# config 4: 1M-drone population sharded across N GPUs (131072 per GPU at N=8)
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --drones-per-gpu $((1048576/N)) --steps 5 --generations 10 > gpurun_out/config4_n$N.json 2> gpurun_out/config4.err
# config 5: 100 generations of the full loop, 1M drones
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 scripts/evolve_b200.py --drones 1048576 --generations 100 --timesteps 2000 2>&1 | grep -v OMP | tail -4 > gpurun_out/config5_n$N.txt
cat gpurun_out/config5_n$N.txt
