#!/bin/bash
# Dense replicate designs: C3d (12,200 rows in groups of 3, sample-side Gram of order 8,134) and
# C5 (60,000 rows in groups of 3, gene-side 20,000^2 Gram).  usage: gpurun -- 'bash scripts/gpu_dense.sh <tag> [N]'
TAG=${1:-r01dense}
N=${2:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for w in ${WORKLOADS:-C3d C5}; do
  if [ $N = 1 ]; then
    timeout 900 python bench.py --workload $w --steps ${STEPS:-5} --warmup 3 --no-e2e > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --workload $w --gpus $N --steps ${STEPS:-5} --warmup 3 --no-e2e > $OUT/bench_${w}_n$N.json 2> $OUT/bench_${w}_n$N.err
  fi
  echo "$w exit $?"; tail -c 2500 $OUT/bench_${w}*.json; tail -3 $OUT/bench_${w}*.err
done
