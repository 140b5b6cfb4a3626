#!/bin/bash
# Round 2 (second session): segmented landing of the slab (per-segment mbarriers) against the shipped kernel.
out=gpurun_out/r02b_seg.jsonl; : > $out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "spmm" > gpurun_out/r02b_spmm_tests.log 2>&1; tail -3 gpurun_out/r02b_spmm_tests.log
for v in base main seg2 seg4 seg8x4 seg8x6; do
  echo "{\"variant\": \"$v\"}" >> $out
  if [ $v == main ]; then unset GCS_LIB_PATH; else export GCS_LIB_PATH=$PWD/gcn-string_b200/variants/lib$v.so; fi
  timeout 300 python scripts/spmm_bench.py --mode slab4 --both --check --iters 30 --ldy 1280 >> $out 2>>gpurun_out/r02b_seg.err
done
unset GCS_LIB_PATH
run() { timeout 400 python bench.py --steps 8 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],3), round(d['ops']['spmm_bwd']['ms_per_call']*1e3,1))"; }
run main
GCS_LIB_PATH=$PWD/gcn-string_b200/variants/libbase.so run base
run main2
cat $out | cut -c1-260
tail -3 gpurun_out/r02b_seg.err
