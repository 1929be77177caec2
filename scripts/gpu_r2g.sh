#!/bin/bash
TAG=${1:-r2g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/host.txt; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA" >> $OUT/host.txt; free -g | head -2 >> $OUT/host.txt
cat $OUT/host.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu exit $?" | tee -a $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
e2e() {  # name, env...
  local name=$1; shift
  env "$@" RUVIII_PIPE_TRACE=1 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "bench $name ($*) exit $?"
  grep "stage trace" $OUT/bench_$name.err | tail -1
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, "parity", d.get("parity_rel_frob"))
    print("   e2e pinned %.1f ms, pageable %.1f ms (x%.2f), identical %s %s" % (d["e2e"]["ms_per_step"], d["e2e_pageable"]["ms_per_step"], d["e2e_pageable"]["vs_pinned"], d["e2e"]["bit_identical_to_sequential_path"], d["e2e_pageable"]["bit_identical_to_sequential_path"]))
except Exception as e:
    print("   no result:", e)
PY
}
e2e default RUVIII_X=0
e2e nont RUVIII_NT_COPY=0
e2e t8 RUVIII_HOST_THREADS=8
e2e t32 RUVIII_HOST_THREADS=32
e2e nocache RUVIII_YC_CACHE=0
