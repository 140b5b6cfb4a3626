#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02z_gpu_tests.log 2>&1; tail -2 gpurun_out/r02z_gpu_tests.log
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/r02z_bench_default.json 2> gpurun_out/r02z_bench_default.err; grep -E "Elapsed|Maximum resident" gpurun_out/r02z_bench_default.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r02z_bench_default.json'))
print('default bench: steps', d['steps'], 'warmup', d['warmup'], 'step', round(d['ms_per_step'],3), 'graphs/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'cpu', round(d['cpu_baseline']['value'],1), 'allocs', d['per_step_ms']['device_allocs_in_timed_region'], 'max step', max(d['per_step_ms']['value']), max(d['per_step_ms']['e2e']))
P
