#!/bin/bash
# round 2: persistent eigensolver iteration -- spectra tests, C3 bench with the in-kernel phase trace.
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 300 python -m pytest tests/test_eigen_spectra.py -m gpu -q --timeout 120 > $OUT/pytest_eigen.log 2>&1; echo "pytest eigen exit $?" | tee -a $OUT/pytest_eigen.log
tail -12 $OUT/pytest_eigen.log
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
bench trace RUVIII_EIG_TRACE=1
bench default RUVIII_X=0
for v in $VARIANTS; do bench ${v%%:*} ${v#*:}; done
