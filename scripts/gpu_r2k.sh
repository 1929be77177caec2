#!/bin/bash
OUT=gpurun_out/r2k; mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests/test_ncg.py -m gpu -q --timeout 300 > $OUT/pytest_ncg.log 2>&1; echo "pytest ncg exit $?"; tail -12 $OUT/pytest_ncg.log | cut -c1-300
