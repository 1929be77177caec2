#!/bin/bash
TAG=${1:-r2n}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout -s KILL 60 ./scripts/lat_probe > $OUT/lat_probe.txt 2>&1; echo "lat_probe exit $?"; cat $OUT/lat_probe.txt
timeout -s KILL 600 python -m pytest tests/test_ncg.py -m gpu -q -x --timeout 300 > $OUT/pytest_ncg.log 2>&1; echo "pytest ncg exit $?"; tail -25 $OUT/pytest_ncg.log
