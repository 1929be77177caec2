#!/bin/bash
# Round 2 profile: tests first (the same binaries), then the launch list of two C3 passes and full captures of the main
# kernels.  usage: gpurun --timeout 2400 -- 'bash scripts/gpu_r2_ncu.sh <tag>'
TAG=${1:-r2ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest gpu exit $?" | tee -a $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout -s KILL 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "default bench exit $?"
python -c "
import json,sys
d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1])
print('  ms/step %.3f e2e %.1f pageable %.1f clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pageable']['ms_per_step'], d['clocks']))
"
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra"
timeout -s KILL 300 $CMD > $OUT/plain.log 2>&1 && \
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on \
    -k regex:"update_smem_kernel|gram_dmma_kernel|gram_reduce|helmert_kernel|solve_fused_kernel|project_kernel|ctl_products" -s 12 -c 12 \
    -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"ek_persistent_kernel|ctl_gather_kernel" -c 3 \
    -o $OUT/prof_eig $CMD > $OUT/ncu_eig.log 2>&1
echo "ncu eig exit $?"
ls -la $OUT
for f in prof prof_eig; do
  [ -f $OUT/$f.ncu-rep ] && ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/$f.raw.csv 2> /dev/null
done
du -sh $OUT
