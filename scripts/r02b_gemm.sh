#!/bin/bash
# deferred-epilogue GEMM: unit tests, model parity tests, then step-time A/B against the previous kernel
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "linear or gemm or tensor_core" > gpurun_out/r02b_gemm_tests.log 2>&1; tail -3 gpurun_out/r02b_gemm_tests.log
timeout 1200 python -m pytest tests/test_model_gpu.py -x -q > gpurun_out/r02b_model_tests.log 2>&1; tail -3 gpurun_out/r02b_model_tests.log
run() { timeout 400 python bench.py --steps 8 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); o=d['ops']; print('$1', round(d['ms_per_step'],3), 'fwd', round(o['linear_fwd']['ms_per_call'],4), 'dx', round(o['linear_bwd_input']['ms_per_call'],4), 'dw', round(o['linear_bwd_weight']['ms_per_call'],4), 'spmm', round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],3), 'fwdpass', round(d['fwd_ms_per_step'],3))"; }
run main
GCS_LIB_PATH=$PWD/gcn-string_b200/variants/libgemmold.so run old
run main
GCS_LIB_PATH=$PWD/gcn-string_b200/variants/libgemmold.so run old
