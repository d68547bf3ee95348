#!/bin/bash
# ncu captures of one cfg2 training step (through gpurun, one GPU):
#   TAG_launches.csv  launch list (--metrics gpu__time_duration.sum) of two steps
#   TAG_full.ncu-rep  --set full --import-source on of the five kernels of one step
# Usage: bash tools/ncu_step.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --extras none"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_bench.log 2>&1
# skip the set-up launches and the warm-up steps: capture 5 kernels of a steady-state step
ncu --set full --clock-control none --import-source on -k "regex:k_pipe_tcg|k_pipe_tn|k_finalize" \
    -s 15 -c 5 -o gpurun_out/${TAG}_full -f $BENCH > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/${TAG}_full.ncu-rep
