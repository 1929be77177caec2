#!/bin/bash
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_aux.sh <tag>'   -- tests of the 8f rows + their timings + the eigensolver trace
TAG=${1:-r01aux}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -k "not tma_kernel_variants" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -6 $OUT/pytest.log
timeout 600 python scripts/bench_aux.py > $OUT/bench_aux.json 2> $OUT/bench_aux.err; echo "bench_aux exit $?"; cat $OUT/bench_aux.json; tail -3 $OUT/bench_aux.err
RUVIII_EIG_TRACE=1 RUVIII_PDL=0 timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/trace.json 2> $OUT/trace.err; echo "trace exit $?"
grep "eig trace" $OUT/trace.err | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
for v in ${AUX_VARIANTS:-gather1:RUVIII_GATHER_AT=1}; do
  name=${v%%:*}; kv=${v#*:}
  env $kv timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "variant $name exit $?"
done
python - $OUT <<'PY'
import json, sys, glob, os
for f in sorted(glob.glob(os.path.join(sys.argv[1], "bench*.json"))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if "stage_ms" not in d: continue
        print(os.path.basename(f), "ms/step %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004 and k != "ms_h2d"},
              d.get("e2e") and round(d["e2e"]["ms_per_step"], 1))
    except Exception as e:
        print(f, "ERR", e)
PY
