#!/bin/bash
# ncu captures of round 2 (dev tool; run under gpurun, one GPU): launch lists + one full capture per workload
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --generations 3"
$CMD > gpurun_out/r02c_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02c_launches.csv $CMD > gpurun_out/r02c_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fly_mlp_f64 -s 3 -c 1 -o gpurun_out/r02c_fly_survivors $CMD > gpurun_out/r02c_ncu2.log 2>&1
CMDR="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --generations 0 --population random --drones-per-gpu 262144"
$CMDR > gpurun_out/r02c_plain_random.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02c_launches_random.csv $CMDR > gpurun_out/r02c_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fly_mlp_f64 -s 3 -c 1 -o gpurun_out/r02c_fly_random $CMDR > gpurun_out/r02c_ncu4.log 2>&1
# a GA-offspring generation: the 4th fly launch after the timed region (generations leg: warm-up + 3 generations)
ncu --set full --clock-control none --import-source on -k regex:fly_mlp_f64 -s 11 -c 1 -o gpurun_out/r02c_fly_offspring $CMD > gpurun_out/r02c_ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep
