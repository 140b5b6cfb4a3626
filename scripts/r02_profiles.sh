#!/bin/bash
# Round-2 evidence: ncu full capture of the slab aggregation kernel (fwd with prologue, plain), launch list of the bench command.
set -x
python scripts/spmm_bench.py --mode slab4 --both --iters 5 > gpurun_out/r02_slab4_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:spmm_slab -s 3 -c 1 -o gpurun_out/r02_slab4_fwd -f python scripts/spmm_bench.py --mode slab4 --iters 2 > gpurun_out/r02_ncu_slab4_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmm_slab -s 3 -c 1 -o gpurun_out/r02_slab4_plain -f python scripts/spmm_bench.py --mode slab4 --no-transform --iters 2 > gpurun_out/r02_ncu_slab4_plain.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_steps2.json 2> gpurun_out/r02_bench_steps2.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_bench.csv | tail -4
