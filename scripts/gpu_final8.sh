#!/bin/bash
# Final 8-GPU evidence: gpurun --gpus 8 --timeout 1200 -- 'bash scripts/gpu_final8.sh <tag>'
TAG=${1:-r2z8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpu.csv 2>&1; nproc >> $OUT/gpu.csv
timeout -s KILL 400 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_multi.log
tail -4 $OUT/pytest_multi.log
run() {  # name n extra-args env...
  local name=$1 n=$2 args=$3; shift 3
  env "$@" NCCL_DEBUG=WARN timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline $args > $OUT/scale_$name.json 2> $OUT/scale_$name.err
  echo "run $name exit $?"
  python - "$OUT/scale_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   N=%d ms/step %.3f value %.3e" % (d["n_gpus"], d["ms_per_step"], d["value"]), {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, "parity", d.get("parity_rel_frob"))
    if d.get("e2e"): print("   e2e ms", round(d["e2e"]["ms_per_step"], 1), "pageable", d.get("e2e_pageable") and round(d["e2e_pageable"]["ms_per_step"], 1))
    for c in ("C4", "C5"):
        x = (d.get("configs") or {}).get(c)
        if x: print("  ", c, x.get("ms_per_step"), x.get("parity"))
except Exception as e:
    print("   no result:", e)
PY
}
run 8 8 "" RUVIII_X=0
run 8_nccl 8 "--no-extra --no-e2e" RUVIII_P2P=0
run 8_p2pall 8 "--no-extra --no-e2e" RUVIII_P2P_MIN_KB=0
run 4 4 "--no-extra --no-e2e" RUVIII_X=0
run 2 2 "--no-extra --no-e2e" RUVIII_X=0
