#!/bin/bash
# Round 2: slab aggregation kernel, correctness + tuning sweep (one B200).
set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "spmm" > gpurun_out/r02_spmm_tests.log 2>&1
tail -5 gpurun_out/r02_spmm_tests.log
out=gpurun_out/r02_spmm_explore.jsonl
: > $out
timeout 300 python scripts/spmm_bench.py --mode rb4 --both --iters 20 >> $out 2>gpurun_out/r02_spmm_explore.err
for cfg in "2 0"; do
  set -- $cfg
  echo "{\"stages\": $1, \"stage_bytes\": $2}" >> $out
  timeout 300 python scripts/spmm_bench.py --mode slab2,slab4 --both --check --iters 20 --ldy 1280 --param 10 $1 --param 11 $2 >> $out 2>>gpurun_out/r02_spmm_explore.err
done
cat $out
tail -5 gpurun_out/r02_spmm_explore.err
