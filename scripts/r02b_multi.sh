#!/bin/bash
# usage: r02b_multi.sh N  — NCCL correctness check + bench on N GPUs of one node (+ the one-GPU bench on the same box)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --check > gpurun_out/r02b_check_${N}gpu.json 2> gpurun_out/r02b_check_${N}gpu.err; cut -c1-400 gpurun_out/r02b_check_${N}gpu.json; tail -2 gpurun_out/r02b_check_${N}gpu.err
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_${N}gpu.json 2> gpurun_out/r02b_bench_${N}gpu.err; tail -2 gpurun_out/r02b_bench_${N}gpu.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_1gpu_samebox.json 2> gpurun_out/r02b_bench_1gpu_samebox.err
python - <<PY
import json
for f in ("gpurun_out/r02b_bench_${N}gpu.json", "gpurun_out/r02b_bench_1gpu_samebox.json"):
    try:
        d = json.load(open(f)); ps = d["per_step_ms"]
        print(f, d["n_gpus"], round(d["ms_per_step"], 3), round(d["value"]), round(d["e2e"]["value"]), "max step", max(ps["value"]), max(ps["e2e"]), ps.get("device_allocs_in_timed_region"))
    except Exception as e: print(f, "failed", e)
PY
