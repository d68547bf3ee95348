#!/bin/bash
# ncu --set full captures of the hot kernels of the other BASELINE configs (one GPU, through
# gpurun): cfg3 (large-graph SpMM + tcgen05 transform), cfg4 (fused Kipf / Duvenaud tile
# kernels), cfg5 (power-law rows).  One launch of each kernel, taken after the warm-up.
# Usage: bash tools/ncu_configs.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:k_agg_tc128" -s 2 -c 2 \
    -o gpurun_out/${TAG}_cfg3 -f python tools/bench_configs.py cfg3 > gpurun_out/${TAG}_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_duv_fwd|k_duv_bwd|k_kipf_fwd|k_kipf_bwd" -s 12 -c 6 \
    -o gpurun_out/${TAG}_cfg4 -f python tools/bench_configs.py cfg4 > gpurun_out/${TAG}_cfg4.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_aggregate|k_tc_rows|k_tc_tn" -s 30 -c 8 \
    -o gpurun_out/${TAG}_cfg5 -f python tools/bench_configs.py cfg5 > gpurun_out/${TAG}_cfg5.log 2>&1
ls -la gpurun_out/${TAG}_cfg*.ncu-rep
tail -2 gpurun_out/${TAG}_cfg3.log gpurun_out/${TAG}_cfg4.log gpurun_out/${TAG}_cfg5.log | cut -c1-300
