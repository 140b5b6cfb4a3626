#!/bin/bash
# A/B of kernel-variant libraries: scripts/r02b_ab.sh "variants..." [spmm_bench args]; 'main' = the in-tree library.
out=gpurun_out/r02b_ab.jsonl; : > $out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "spmm" > gpurun_out/r02b_spmm_tests.log 2>&1; tail -3 gpurun_out/r02b_spmm_tests.log
for v in $1; do
  echo "{\"variant\": \"$v\"}" >> $out
  if [ $v == main ]; then unset GCS_LIB_PATH; else export GCS_LIB_PATH=$PWD/gcn-string_b200/variants/lib$v.so; fi
  timeout 300 python scripts/spmm_bench.py --mode slab4 --both --check --iters 30 --ldy 1280 >> $out 2>>gpurun_out/r02b_ab.err
done
unset GCS_LIB_PATH
run() { timeout 400 python bench.py --steps 8 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],3), round(d['ops']['spmm_bwd']['ms_per_call']*1e3,1))"; }
for v in $1 $1; do
  if [ $v == main ]; then unset GCS_LIB_PATH; else export GCS_LIB_PATH=$PWD/gcn-string_b200/variants/lib$v.so; fi
  run $v
done
python - <<'P'
import json
for l in open('gpurun_out/r02b_ab.jsonl'):
    d=json.loads(l)
    print(d.get('variant') or (d['prologue'], d['us'], d['frac_measured_hbm'], d.get('bitwise_equal_rows_kernel')))
P
tail -3 gpurun_out/r02b_ab.err
