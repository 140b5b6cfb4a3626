#!/bin/bash
# model parity tests + the default bench (per-op summary)
timeout 1800 python -m pytest tests/test_model_gpu.py -x -q > gpurun_out/r02b_model_tests.log 2>&1; tail -3 gpurun_out/r02b_model_tests.log
for i in 1 2; do
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02b_bench.json')); o=d['ops']
print('step', round(d['ms_per_step'],3), 'graphs/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'spmm', round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], 'fwd', round(d['fwd_ms_per_step'],3))
print({k: round(v['ms_per_call'],4) for k,v in o.items()})
P
done
tail -3 gpurun_out/r02b_bench.err
