#!/bin/bash
# Build an experiment variant of libnngp_b200.so: only bf_tri_kernel.cu is recompiled with the given -D flags; every other
# object comes from the regular build.   tools/build_tri_variant.sh NAME -DNNGP_POW_FUSED=0 ...   (NNGP_B200_LIB selects it)
set -e
cd "$(dirname "$0")/../spnngp_b200/csrc"
name=$1; shift
mkdir -p ../variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c bf_tri_kernel.cu -o ../variants/tri_$name.o
objs=$(ls *.o | grep -v bf_tri_kernel.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libnngp_$name.so $objs ../variants/tri_$name.o
cuobjdump -res-usage ../variants/tri_$name.o | grep -A1 "bf_tri_kernelILi4E" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | paste - -
