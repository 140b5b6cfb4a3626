#!/bin/bash
# usage: r02_multi_full.sh N [full]  — N GPUs of one node: default bench; with "full" also --check, cfg5 and the 100k-graph epoch
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { name=$1; shift; timeout 900 $TR bench.py --gpus $N "$@" > gpurun_out/r02_${name}_${N}gpu.json 2> gpurun_out/r02_${name}_${N}gpu.err; tail -1 gpurun_out/r02_${name}_${N}gpu.err | cut -c1-200; }
run bench --steps 20 --warmup 3
if [ "$2" = "full" ]; then
  run check --check
  run bench_cfg5 --workload cfg5 --steps 10 --warmup 3
  run epoch100k --epoch 100000
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_*_${N}gpu.json")):
    try:
        d = json.load(open(f))
        if "check" in d: print(f, d)
        else: print(f, d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "graphs/s", round(d["value"]), "e2e", round(d.get("e2e", {}).get("value", 0)), d["config"]["workload"][:40])
    except Exception as e: print(f, "failed", e)
PY
