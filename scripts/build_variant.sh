#!/bin/bash
# Build a kernel-variant library for A/B runs on the GPU box:
#   scripts/build_variant.sh NAME spmm_slab [-DFLAG ...] [SRC=path.cu]
# compiles csrc/spmm_slab.cu (or SRC) with the extra flags in place of the regular spmm_slab.o and links
# gcn-string_b200/variants/libNAME.so (select it with GCS_LIB_PATH); all other objects come from the regular build.
set -e
name=$1; unit=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
cs=$root/gcn-string_b200/csrc
src=$cs/$unit.cu
flags=()
for a in "$@"; do case $a in SRC=*) src=${a#SRC=};; *) flags+=("$a");; esac; done
mkdir -p $root/gcn-string_b200/variants $cs/build/var_$name
obj=$cs/build/var_$name/$unit.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "${flags[@]}" -I$cs -c $src -o $obj
objs=""
for o in $cs/build/*.o; do
  if [ "$(basename $o)" == "$unit.o" ]; then objs="$objs $obj"; else objs="$objs $o"; fi
done
/usr/local/cuda/bin/nvcc -shared -o $root/gcn-string_b200/variants/lib$name.so $objs -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lrt -lpthread
echo $root/gcn-string_b200/variants/lib$name.so
