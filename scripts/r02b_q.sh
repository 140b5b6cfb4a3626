#!/bin/bash
out=gpurun_out/r02b_q.jsonl; : > $out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "spmm" > gpurun_out/r02b_spmm_tests.log 2>&1; tail -2 gpurun_out/r02b_spmm_tests.log
b() { echo "{\"variant\": \"$1\"}" >> $out; shift; timeout 300 python scripts/spmm_bench.py --both --check --iters 30 --ldy 1280 "$@" >> $out 2>>gpurun_out/r02b_q.err; }
GCS_LIB_PATH=$PWD/gcn-string_b200/variants/libdyn.so b dyn_deg12 --mode slab4
b main_deg12 --mode slab4
b main_deg32 --mode slab4 --deg 32
python - <<'P'
import json
for l in open('gpurun_out/r02b_q.jsonl'):
    d=json.loads(l)
    print(d.get('variant') or (d['mode'], d['deg'], d['prologue'], d['us'], d['frac_measured_hbm'], d.get('max_rel_diff_vs_rows_kernel')))
P
tail -3 gpurun_out/r02b_q.err
