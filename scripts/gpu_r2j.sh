#!/bin/bash
TAG=${1:-r2j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu exit $?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
RUVIII_PIPE_TRACE=1 timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-extra > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
grep "stage trace" $OUT/bench.err | tail -2
python - "$OUT/bench.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("   ms/step %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, "parity", d.get("parity_rel_frob"))
print("   e2e pinned %.1f ms, pageable %.1f ms (x%.2f), identical %s %s" % (d["e2e"]["ms_per_step"], d["e2e_pageable"]["ms_per_step"], d["e2e_pageable"]["vs_pinned"], d["e2e"]["bit_identical_to_sequential_path"], d["e2e_pageable"]["bit_identical_to_sequential_path"]))
print("   cpu", d["cpu_baseline"])
PY
python scripts/bench_aux.py > $OUT/bench_aux.json 2> $OUT/bench_aux.err; echo "aux exit $?"; tail -c 1500 $OUT/bench_aux.json
