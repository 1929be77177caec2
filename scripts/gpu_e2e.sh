#!/bin/bash
# usage: gpurun -- 'bash scripts/gpu_e2e.sh <tag>'  -- end-to-end (host to host) leg of bench.py under different gather splits
TAG=${1:-r01e2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -k "pipelined or golden or stacked" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest.log
for pct in ${PCTS:-100 67 50 33 0}; do
  RUVIII_ZC_PCT=$pct timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_zc$pct.json 2> $OUT/bench_zc$pct.err
  python - $OUT/bench_zc$pct.json $pct <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("zc_pct", sys.argv[2], "e2e ms %.2f" % d["e2e"]["ms_per_step"], "pipelined", d["e2e"]["pipelined"], "resident ms %.3f" % d["ms_per_step"])
PY
done
RUVIII_PIPELINE=2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_nopipe.json 2> $OUT/bench_nopipe.err
python - $OUT/bench_nopipe.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("no pipeline: e2e ms %.2f" % d["e2e"]["ms_per_step"], "pipelined", d["e2e"]["pipelined"])
PY
