#!/bin/bash
# Multi-GPU check, round 2: gpurun --gpus N --timeout 1800 -- 'bash scripts/gpu_multi2.sh <tag> <N> [skip-tests]'
TAG=${1:-r2m}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpu.csv 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
nproc > $OUT/host.txt
if [ -z "$3" ]; then
timeout -s KILL 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_multi.log
tail -8 $OUT/pytest_multi.log
fi
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n = 1 ]; then
    timeout -s KILL 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  else
    NCCL_DEBUG=${NCCL_DEBUG:-WARN} timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --probes > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  fi
  echo "scale $n exit $?"; tail -3 $OUT/scale_$n.err
  python - "$OUT/scale_$n.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   N=%d ms/step %.3f value %.3e" % (d["n_gpus"], d["ms_per_step"], d["value"]), {k: round(v, 3) for k, v in d["stage_ms"].items() if v > 0.004}, "parity", d.get("parity_rel_frob"))
    print("   e2e pinned %.1f ms, pageable %.1f ms, identical %s %s" % (d["e2e"]["ms_per_step"], d["e2e_pageable"]["ms_per_step"], d["e2e"]["bit_identical_to_sequential_path"], d["e2e_pageable"]["bit_identical_to_sequential_path"]))
    for name, c in (d.get("configs") or {}).items():
        print("  ", name, c.get("ms_per_step"), (c.get("e2e") or {}).get("ms_per_step"), c.get("parity"), c.get("error"))
    if d.get("nccl_allreduce_us"): print("   allreduce us", d["nccl_allreduce_us"])
except Exception as e:
    print("   no result:", e)
PY
done
