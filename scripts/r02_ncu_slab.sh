#!/bin/bash
set -x
python scripts/spmm_bench.py --mode slab2 --iters 2 > gpurun_out/r02_slab2_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:spmm_slab -s 3 -c 1 -o gpurun_out/r02_slab2p -f python scripts/spmm_bench.py --mode slab2 --iters 2 > gpurun_out/r02_ncu_slab2p.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
