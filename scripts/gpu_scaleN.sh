#!/bin/bash
# one resident-pass bench line at N GPUs (no extras): gpurun --gpus N --timeout 600 -- 'bash scripts/gpu_scaleN.sh <tag> N'
TAG=${1:-r2sc}; N=${2:-4}
OUT=gpurun_out/$TAG; mkdir -p $OUT
NCCL_DEBUG=WARN timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extra --no-e2e > $OUT/scale_$N.json 2> $OUT/scale_$N.err
echo "run exit $?"
python - "$OUT/scale_$N.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("   N=%d ms/step %.3f value %.3e" % (d["n_gpus"], d["ms_per_step"], d["value"]), {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, "parity", d.get("parity_rel_frob"))
PY
