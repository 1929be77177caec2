#!/bin/bash
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
bench() {  # name, env...
  local name=$1; shift
  env "$@" timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "bench $name ($*) exit $?"
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.3f  " % d["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, d["config"].get("eig_iterations"), "parity", d.get("parity_rel_frob"))
except Exception as e:
    print("   no result:", e)
PY
}
for v in $VARIANTS; do bench ${v%%:*} $(echo ${v#*:} | tr ',' ' '); done
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "staged or pipelined" > $OUT/pytest_staged.log 2>&1; echo "pytest staged exit $?" | tee -a $OUT/pytest_staged.log
tail -15 $OUT/pytest_staged.log
RUVIII_PIPE_TRACE=1 timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_full.json 2> $OUT/bench_full.err; echo "bench full exit $?"
grep "stage trace" $OUT/bench_full.err | tail -4
tail -3 $OUT/bench_full.err
python - "$OUT/bench_full.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("ms_per_step", "parity_rel_frob", "e2e", "e2e_pageable", "probes"):
    print(k, d.get(k))
print("stage", {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004})
print("kernels", {k: (round(v["achieved"], 1), round(v["frac"], 3)) for k, v in d["kernels"].items()})
for name, c in (d.get("configs") or {}).items():
    print(name, {k: v for k, v in c.items() if k not in ("workload",)})
PY
