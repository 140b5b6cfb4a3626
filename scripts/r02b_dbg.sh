#!/bin/bash
GCS_BENCH_DEFAULT_ALLOCATOR=1 GCS_LOADER_DEBUG=1 timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02b_dbg.json 2> gpurun_out/r02b_dbg_1.err
grep "model" gpurun_out/r02b_dbg_1.err | head -18
python -c "
import json; d=json.load(open('gpurun_out/r02b_dbg.json')); print(d['per_step_ms']['value_host_enqueue'])"
