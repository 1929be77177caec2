#!/bin/bash
# usage: gpurun -- 'bash scripts/gpu_ncu_aux.sh <tag>'  -- ncu full capture of the 8f kernels (K0, K10, K11) under bench_aux.py
TAG=${1:-r01ncuaux}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CMD="python scripts/bench_aux.py --steps 1"
timeout 300 $CMD > $OUT/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"center_kernel|colsum_kernel|gene_stats_kernel|group_mean_kernel|helmert_kernel" -c 10 \
    -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
echo "ncu exit $?"; ls -la $OUT
