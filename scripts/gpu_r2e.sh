#!/bin/bash
TAG=${1:-r2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
bench() {  # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "bench $name ($*) exit $?"
  grep "eig persist" $OUT/bench_$name.err | tail -1 | cut -c1-2500
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.3f  " % d["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, d["config"].get("eigensolver"), d["config"].get("eig_iterations"))
except Exception as e:
    print("   no result:", e)
PY
}
for v in $VARIANTS; do bench ${v%%:*} $(echo ${v#*:} | tr ',' ' '); done
