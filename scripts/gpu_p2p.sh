#!/bin/bash
# P2P allreduce check on N GPUs: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_p2p.sh <tag> <N>'
TAG=${1:-r2p}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {  # name env...
  local name=$1; shift
  env "$@" NCCL_DEBUG=WARN timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extra $BENCH_ARGS > $OUT/scale_$name.json 2> $OUT/scale_$name.err
  echo "run $name exit $?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/scale_$name.err | tail -3
  python - "$OUT/scale_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   N=%d ms/step %.3f value %.3e" % (d["n_gpus"], d["ms_per_step"], d["value"]), {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, "parity", d.get("parity_rel_frob"))
    print("   e2e pinned %.1f ms, pageable %.1f ms, identical %s %s" % (d["e2e"]["ms_per_step"], d["e2e_pageable"]["ms_per_step"], d["e2e"]["bit_identical_to_sequential_path"], d["e2e_pageable"]["bit_identical_to_sequential_path"]))
except Exception as e:
    print("   no result:", e)
PY
}
run p2p RUVIII_P2P=1
for c in $P2P_CTAS; do run p2p_c$c RUVIII_P2P=1 RUVIII_P2P_CTAS=$c; done
run nccl RUVIII_P2P=0
timeout -s KILL 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 -k "torchrun" > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_multi.log
