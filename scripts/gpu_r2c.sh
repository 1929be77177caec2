#!/bin/bash
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for rb in 8 16 32; do
RUVIII_EIG_RB=$rb timeout -s KILL 200 python -m pytest tests/test_eigen_spectra.py -m gpu -q --timeout 120 -k persistent_matches 2>&1 | grep -E "AssertionError|passed|failed" | head -5
done
