#!/bin/bash
# 8-GPU check: gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_multi8.sh <tag>'
TAG=${1:-r01m8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpu.csv 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_multi.log
tail -6 $OUT/pytest_multi.log
for n in 8 4 2 1; do
  if [ $n = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  fi
  echo "scale $n exit $?"
  grep "^{" $OUT/scale_$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  N=%d value %.3e ms %.3f e2e %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'] and (round(d['e2e']['value']/1e9,2), round(d['e2e']['ms_per_step'],1))))
print('  ', {k: round(v,3) for k,v in d['stage_ms'].items() if v > 0.004})
"
  tail -2 $OUT/scale_$n.err | cut -c1-200
done
