#!/bin/bash
# Final 1-GPU evidence run: the GPU tests, smoke(), the default bench line (with extras), the reference arm, the §8f
# rows, the in-kernel eigensolver trace and the full-size CPU reference.  usage: gpurun --timeout 2400 -- 'bash scripts/gpu_final1.sh TAG'
TAG=${1:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/host.txt; free -g | head -2 >> $OUT/host.txt; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv >> $OUT/host.txt
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu exit $?" | tee -a $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
timeout -s KILL 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - "$OUT/bench.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.3f value %.3e" % (d["ms_per_step"], d["value"]), {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004})
    print("   e2e", d["e2e"], "pageable", d.get("e2e_pageable"))
    print("   roofline", d["roofline"]); print("   cpu", d["cpu_baseline"]); print("   clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
    for c in ("C4", "C5"):
        x = d.get("configs", {}).get(c, {}); print("  ", c, {k: x.get(k) for k in ("ms_per_step", "e2e_ms", "parity")})
except Exception as e:
    print("   no result:", e)
PY
timeout -s KILL 400 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm exit $?"; cut -c1-600 $OUT/bench_reference.json
RUVIII_EIG_TRACE=1 timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $OUT/bench_trace.json 2> $OUT/bench_trace.err; echo "trace exit $?"
timeout -s KILL 400 python scripts/bench_aux.py > $OUT/aux.json 2> $OUT/aux.err; echo "aux exit $?"; cut -c1-700 $OUT/aux.json
if [ "$2" = "full-ref" ]; then
( time timeout -s KILL 1200 python bench.py --impl reference --ref-full --workload C3 --steps 1 --warmup 0 ) > $OUT/reference_full_C3.json 2> $OUT/reference_full_C3.err
echo "ref full exit $?"; cut -c1-900 $OUT/reference_full_C3.json; tail -4 $OUT/reference_full_C3.err
fi
