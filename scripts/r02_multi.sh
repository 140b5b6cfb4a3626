#!/bin/bash
# usage: r02_multi.sh N  — multi-GPU correctness + short bench on N GPUs of one node
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --check > gpurun_out/r02_check_${N}gpu.json 2> gpurun_out/r02_check_${N}gpu.err; cat gpurun_out/r02_check_${N}gpu.json; tail -2 gpurun_out/r02_check_${N}gpu.err
timeout 300 $TR bench.py --gpus $N --check --sync-bn > gpurun_out/r02_check_syncbn_${N}gpu.json 2> gpurun_out/r02_check_syncbn_${N}gpu.err; cat gpurun_out/r02_check_syncbn_${N}gpu.json; tail -2 gpurun_out/r02_check_syncbn_${N}gpu.err
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; tail -2 gpurun_out/r02_bench_${N}gpu.err
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --torch-allreduce > gpurun_out/r02_bench_${N}gpu_torchar.json 2> gpurun_out/r02_bench_${N}gpu_torchar.err; tail -2 gpurun_out/r02_bench_${N}gpu_torchar.err
python - <<PY
import json
for f in ("gpurun_out/r02_bench_${N}gpu.json", "gpurun_out/r02_bench_${N}gpu_torchar.json"):
    try:
        d = json.load(open(f)); print(f, d["n_gpus"], round(d["ms_per_step"], 3), round(d["value"]), round(d["e2e"]["value"]), d["config"].get("collective"))
    except Exception as e: print(f, "failed", e)
PY
