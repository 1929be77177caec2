#!/bin/bash
TAG=${1:-r2u}
OUT=gpurun_out/$TAG; mkdir -p $OUT
RUVIII_PIPE_TRACE=1 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
grep "stage trace" $OUT/bench.err | tail -12
python - "$OUT/bench.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("   ms/step %.3f e2e %.1f ms pageable %.1f ms" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e_pageable"]["ms_per_step"]))
PY
