#!/bin/bash
run() { timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02b_stall.json 2> gpurun_out/r02b_stall.err; python - <<P
import json
d=json.load(open('gpurun_out/r02b_stall.json'))
ps=d['per_step_ms']
bad = max(ps['value'])>20 or max(ps['e2e'])>20
print('$1 step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'max', max(ps['value']), max(ps['e2e']), 'host max', max(ps['value_host_enqueue']), 'allocs', ps['device_allocs_in_timed_region'])
if bad: print(ps)
P
}
for i in 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15 16; do run r_$i; done
tail -2 gpurun_out/r02b_stall.err
