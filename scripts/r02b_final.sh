#!/bin/bash
# End-of-round evidence on one B200: GPU suite, smoke, bench (+ cpu baseline), launch list, ncu captures of the aggregation kernel,
# cfg4 sweep, cfg5, the 100k-graph epoch.  Every ncu run follows the same command exiting 0 without ncu.
set -x
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gpu_tests.log 2>&1; tail -2 gpurun_out/r02f_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/r02f_smoke.log 2>&1; tail -3 gpurun_out/r02f_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -2 gpurun_out/r02f_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_steps2.json 2> gpurun_out/r02f_bench_steps2.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_bench.log 2>&1
timeout 300 python scripts/spmm_bench.py --mode slab4 --both --iters 20 --ldy 1280 --check > gpurun_out/r02f_slab4.jsonl 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_slab -s 3 -c 1 -o gpurun_out/r02f_slab4_fwd -f python scripts/spmm_bench.py --mode slab4 --iters 2 --ldy 1280 > gpurun_out/r02f_ncu_slab4_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_slab -s 3 -c 1 -o gpurun_out/r02f_slab4_plain -f python scripts/spmm_bench.py --mode slab4 --no-transform --iters 2 --ldy 1280 > gpurun_out/r02f_ncu_slab4_plain.log 2>&1
timeout 900 python scripts/spmm_bench.py --sweep --mode rb4,slab4 --iters 10 --check > gpurun_out/r02f_cfg4_sweep.jsonl 2> gpurun_out/r02f_cfg4_sweep.err
timeout 600 python bench.py --workload cfg5 --steps 10 --no-cpu-baseline > gpurun_out/r02f_bench_cfg5.json 2> gpurun_out/r02f_bench_cfg5.err
timeout 600 python bench.py --epoch 100000 > gpurun_out/r02f_epoch100k.json 2> gpurun_out/r02f_epoch100k.err
ls -la gpurun_out/r02f_* | tail -20
