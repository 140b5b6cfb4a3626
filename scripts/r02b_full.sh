#!/bin/bash
out=gpurun_out/r02b_full.jsonl; : > $out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "spmm" > gpurun_out/r02b_spmm_tests.log 2>&1; tail -5 gpurun_out/r02b_spmm_tests.log
b() { echo "{\"variant\": \"$1\"}" >> $out; shift; timeout 300 python scripts/spmm_bench.py --both --check --iters 30 --ldy 1280 "$@" >> $out 2>>gpurun_out/r02b_full.err; }
GCS_LIB_PATH=$PWD/gcn-string_b200/variants/libdyn.so b dyn_deg12 --mode slab4
b full_deg12 --mode slab4
GCS_LIB_PATH=$PWD/gcn-string_b200/variants/libdyn.so b dyn_deg32 --mode slab4 --deg 32
b full_deg32 --mode slab4,rb4 --deg 32
b full_deg16 --mode slab4 --deg 16
b full_deg64 --mode slab4 --deg 64
b cfg5shape --mode rb4 --graphs 64 --n-mean 5000 --deg 32 --hidden 512 --ldy 2560
python - <<'P'
import json
for l in open('gpurun_out/r02b_full.jsonl'):
    d=json.loads(l)
    print(d.get('variant') or (d['mode'], d['deg'], d['prologue'], d['us'], d['frac_measured_hbm'], d.get('max_rel_diff_vs_rows_kernel')))
P
tail -3 gpurun_out/r02b_full.err
run() { timeout 400 python bench.py --steps 8 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],3), round(d['ops']['spmm_bwd']['ms_per_call']*1e3,1))"; }
run main
GCS_LIB_PATH=$PWD/gcn-string_b200/variants/libdyn.so run dyn
run main
