#!/bin/bash
# full GPU suite + the default bench + cfg5, after a change
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.log 2>&1; tail -3 gpurun_out/r02b_gpu_tests.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02b_bench.json')); o=d['ops']
print('step', round(d['ms_per_step'],3), 'graphs/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'spmm', round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],3), 'clocks', d['clocks'])
print({k: round(v['ms_per_call'],4) for k,v in o.items()})
P
timeout 600 python bench.py --workload cfg5 --steps 5 --no-cpu-baseline > gpurun_out/r02b_cfg5.json 2> gpurun_out/r02b_cfg5.err; python -c "
import json; d=json.load(open('gpurun_out/r02b_cfg5.json')); print('cfg5 step', round(d['ms_per_step'],2), 'spmm', round(d['roofline']['us_per_launch'],1), round(d['roofline']['frac'],3))"
