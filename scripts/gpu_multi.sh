#!/bin/bash
# Multi-GPU check: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi.sh <tag> <N>'
TAG=${1:-r01m}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpu.csv 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_multi.log
tail -15 $OUT/pytest_multi.log
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  else
    NCCL_DEBUG=${NCCL_DEBUG:-WARN} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --probes > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  fi
  echo "scale $n exit $?"; tail -c 1500 $OUT/scale_$n.json; echo; tail -3 $OUT/scale_$n.err
done
