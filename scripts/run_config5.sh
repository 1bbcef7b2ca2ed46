#!/bin/bash
# config 4/5 on N GPUs with the peer-memory GA step (dev tool; run under gpurun --gpus N)
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29660 bench.py --gpus $N --steps 10 --generations 10 > gpurun_out/p2p_scale_n$N.json 2> gpurun_out/p2p_scale.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus $N --drones-per-gpu $((1048576/N)) --steps 5 --generations 10 > gpurun_out/p2p_config4_n$N.json 2> gpurun_out/p2p_config4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29662 scripts/evolve_b200.py --drones 1048576 --generations 100 --timesteps 2000 2>&1 | grep -v OMP | tail -3 > gpurun_out/p2p_config5_n$N.txt
cat gpurun_out/p2p_config5_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29663 scripts/p2p_evolution_check.py 2>&1 | grep "p2p =="
