#!/bin/bash
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus $N --check > gpurun_out/r02f_check_${N}gpu.json 2> gpurun_out/r02f_check_${N}gpu.err; cut -c1-300 gpurun_out/r02f_check_${N}gpu.json
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_${N}gpu.json 2> gpurun_out/r02f_bench_${N}gpu.err; tail -1 gpurun_out/r02f_bench_${N}gpu.err
timeout 300 $TR bench.py --gpus $N --epoch 100000 > gpurun_out/r02f_epoch100k_${N}gpu.json 2> gpurun_out/r02f_epoch100k_${N}gpu.err
timeout 300 $TR bench.py --gpus $N --workload cfg5 --steps 5 --no-cpu-baseline > gpurun_out/r02f_bench_cfg5_${N}gpu.json 2> gpurun_out/r02f_bench_cfg5_${N}gpu.err
python - <<PY
import json
for f in ("gpurun_out/r02f_bench_${N}gpu.json", "gpurun_out/r02f_bench_cfg5_${N}gpu.json"):
    try:
        d = json.load(open(f)); ps = d["per_step_ms"]
        print(f, d["n_gpus"], round(d["ms_per_step"], 3), round(d["value"]), round(d["e2e"]["value"]), "max step", max(ps["value"]), max(ps["e2e"]), ps.get("device_allocs_in_timed_region"))
    except Exception as e: print(f, "failed", e)
try:
    e = json.load(open("gpurun_out/r02f_epoch100k_${N}gpu.json")); print("epoch", round(e["value"]), e["config"].get("epoch_ms"))
except Exception as ex: print("epoch failed", ex)
PY
