#!/bin/bash
# Build an experiment variant of libnngp_b200.so: only bf_ws_kernel.cu is recompiled (MC = 15 only) with the given
# -D flags; every other object comes from the regular build.   tools/build_variant.sh NAME -DNNGP_MATERN_BATCH=6 ...
set -e
cd "$(dirname "$0")/../spnngp_b200/csrc"
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DNNGP_WS_ONLY15 "$@" \
  -c bf_ws_kernel.cu -o ../variants/ws_$name.o
objs=$(ls *.o | grep -v bf_ws_kernel.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libnngp_$name.so $objs ../variants/ws_$name.o
cuobjdump -res-usage ../variants/ws_$name.o | grep -A1 "ILi15ELi[04]ELi3" | grep REG
