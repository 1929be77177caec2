#!/bin/bash
TAG=${1:-r2l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu exit $?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
bench() {  # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "bench $name ($*) exit $?"
  grep "eig persist" $OUT/bench_$name.err | tail -1 | cut -c1-1800
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.3f  " % d["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, d["config"].get("eig_iterations"), "parity", d.get("parity_rel_frob"))
except Exception as e:
    print("   no result:", e)
PY
}
bench default RUVIII_X=0
bench trace RUVIII_EIG_TRACE=1
bench chol64 RUVIII_EIG_CHOL32=0
bench jac64 RUVIII_EIG_JAC32=0
bench both64 RUVIII_EIG_CHOL32=0 RUVIII_EIG_JAC32=0
