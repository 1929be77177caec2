#!/bin/bash
# round 2, first GPU call: persistent eigensolver -- spectra tests, C3 bench with the in-kernel phase trace, gather variants.
# usage: gpurun --timeout 900 -- 'bash scripts/gpu_r2a.sh [tag]'
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
timeout -s KILL 300 python -m pytest tests/test_eigen_spectra.py -m gpu -q -x --timeout 120 > $OUT/pytest_eigen.log 2>&1; echo "pytest eigen exit $?" | tee -a $OUT/pytest_eigen.log
tail -15 $OUT/pytest_eigen.log
bench() {  # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "bench $name ($*) exit $?"
  grep "eig persist" $OUT/bench_$name.err | tail -1 | cut -c1-1500
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
bench nopersist RUVIII_EIG_PERSIST=0
bench nocoop RUVIII_EIG_COOP=0
bench gran32 RUVIII_L2_GRAN=32
bench gran64 RUVIII_L2_GRAN=64
bench gather1 RUVIII_GATHER_AT=1
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 > $OUT/pytest_parity.log 2>&1; echo "pytest parity exit $?" | tee -a $OUT/pytest_parity.log
tail -5 $OUT/pytest_parity.log
