#!/bin/bash
TAG=${1:-r2p8}; N=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() {
  local name=$1; shift
  env "$@" NCCL_DEBUG=WARN timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-e2e > $OUT/scale_$name.json 2> $OUT/scale_$name.err
  echo "run $name exit $?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/scale_$name.err | tail -2
  python - "$OUT/scale_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   N=%d ms/step %.3f value %.3e" % (d["n_gpus"], d["ms_per_step"], d["value"]), {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, "parity", d.get("parity_rel_frob"))
except Exception as e:
    print("   no result:", e)
PY
}
run p2p RUVIII_P2P=1
run nccl RUVIII_P2P=0
