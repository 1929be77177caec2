#!/bin/bash
# The literal restatement of ruvIII.R:93-146 at the FULL C3 size on the GPU box's host cores (one step, no warm-up):
# the number that replaces "the bounded sample flatters the CPU".  usage: gpurun --timeout 1800 -- 'bash scripts/gpu_ref_full.sh'
OUT=gpurun_out/r2ref; mkdir -p $OUT
nproc > $OUT/host.txt; free -g | head -2 >> $OUT/host.txt
( time timeout -s KILL 1500 python bench.py --impl reference --ref-full --workload C3 --steps 1 --warmup 0 ) > $OUT/reference_full_C3.json 2> $OUT/reference_full_C3.err
echo "exit $?"; cat $OUT/reference_full_C3.json | cut -c1-900; tail -4 $OUT/reference_full_C3.err
timeout -s KILL 300 python bench.py --impl reference --workload C3 --steps 2 --warmup 1 > $OUT/reference_sample_C3.json 2>> $OUT/reference_full_C3.err; cat $OUT/reference_sample_C3.json | cut -c1-500
