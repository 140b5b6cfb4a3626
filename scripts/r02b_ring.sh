#!/bin/bash
out=gpurun_out/r02b_ring.jsonl; : > $out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "spmm" > gpurun_out/r02b_spmm_tests.log 2>&1; tail -3 gpurun_out/r02b_spmm_tests.log
b() { echo "{\"variant\": \"$1\"}" >> $out; shift; timeout 300 python scripts/spmm_bench.py --mode slab4 --both --check --iters 30 --ldy 1280 "$@" >> $out 2>>gpurun_out/r02b_ring.err; }
GCS_LIB_PATH=$PWD/gcn-string_b200/variants/libdyn.so b dyn
b ring4
b ring3 --param 10 3
b ring2 --param 10 2
b ring4_max96k --param 11 98304
run() { timeout 400 python bench.py --steps 8 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3), round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],3), round(d['ops']['spmm_bwd']['ms_per_call']*1e3,1))"; }
run main
GCS_LIB_PATH=$PWD/gcn-string_b200/variants/libdyn.so run dyn
run main
GCS_LIB_PATH=$PWD/gcn-string_b200/variants/libdyn.so run dyn
python - <<'P'
import json
for l in open('gpurun_out/r02b_ring.jsonl'):
    d=json.loads(l)
    print(d.get('variant') or (d['prologue'], d['us'], d['frac_measured_hbm'], d.get('bitwise_equal_rows_kernel')))
P
tail -3 gpurun_out/r02b_ring.err
