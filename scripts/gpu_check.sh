#!/bin/bash
# One gpurun call: tests, bench, launch list, one full ncu capture.  Everything lands in gpurun_out/.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --probes > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json
RUVIII_GRAM_IMPL=2 RUVIII_UPD_IMPL=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_tma.json 2> $OUT/bench_tma.err; echo "bench tma exit $?"
cat $OUT/bench_tma.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cat $OUT/bench_ref.json
# launch list + full capture of our kernels (only after the same command exited 0 without ncu)
export RUVIII_GRAM_IMPL=${NCU_GRAM_IMPL:-1} RUVIII_UPD_IMPL=${NCU_UPD_IMPL:-1}
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > $OUT/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_list.log 2>&1
timeout 300 $CMD > $OUT/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"update_kernel|update_tma_kernel|gram_dmma_kernel|gram_tma_kernel|ek_gemm|ek_ortho|ek_rr|ctl_products|ac_gram" -s ${NCU_SKIP:-40} -c ${NCU_COUNT:-12} \
    -o $OUT/prof $CMD > $OUT/ncu_full.log 2>&1
ls -la $OUT
