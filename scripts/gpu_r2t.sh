#!/bin/bash
TAG=${1:-r2t}
OUT=gpurun_out/$TAG; mkdir -p $OUT
RUVIII_EIG_TRACE=1 timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_trace.json 2> $OUT/bench_trace.err; echo "trace exit $?"
grep "eig persist" $OUT/bench_trace.err | tail -1 | cut -c1-2500
timeout -s KILL 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - "$OUT/bench.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("   ms/step %.3f  " % d["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, "parity", d.get("parity_rel_frob"))
PY
timeout -s KILL 400 python -m pytest tests/test_eigen_spectra.py -m gpu -q -x --timeout 300 2>&1 | tail -2
