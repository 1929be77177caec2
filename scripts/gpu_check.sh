#!/bin/bash
# One gpurun call: tests, bench (default + kernel variants), launch list, one full ncu capture.  Results in gpurun_out/<tag>.
# usage: gpurun --timeout 2400 -- 'bash scripts/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
tail -2 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q -k "not tma_kernel_variants" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
timeout 300 python -m pytest tests -m gpu -q -k "tma_kernel_variants" --timeout 120 --timeout-method thread > $OUT/pytest_tma.log 2>&1; echo "pytest tma exit $?" >> $OUT/pytest_tma.log
tail -4 $OUT/pytest_tma.log
timeout 600 python bench.py --steps 10 --warmup 3 --probes > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json
variant() {  # name, env...
  local name=$1; shift
  env "$@" timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  echo "variant $name ($*) exit $?"
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.3f  " % d["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004})
except Exception as e:
    print("   no result:", e)
PY
}
VARIANTS=${VARIANTS:-gramtma:RUVIII_GRAM_IMPL=2 updtma:RUVIII_UPD_IMPL=2 updsmem:RUVIII_UPD_IMPL=3}
for v in $VARIANTS; do
  variant ${v%%:*} ${v#*:}
done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cat $OUT/bench_ref.json
# launch list + full capture of our kernels (only after the same command exited 0 without ncu)
for kv in $NCU_ENV; do export $kv; done
echo "ncu with NCU_ENV=$NCU_ENV" > $OUT/ncu_env.txt
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > $OUT/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_list.log 2>&1
timeout 300 $CMD > $OUT/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"${NCU_KERNELS:-update_kernel|update_tma_kernel|update_smem_kernel|gram_dmma_kernel|gram_tma_kernel|helmert_kernel|project_kernel|ctl_products}" -s ${NCU_SKIP:-12} -c ${NCU_COUNT:-6} \
    -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
ls -la $OUT
