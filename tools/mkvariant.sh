#!/bin/bash
# development: tools/mkvariant.sh <name> <file.cu> [-DX=1 ...]  ->  hip-bvh-construction_b200/variants/libb2bvh_<name>.so
# (one source recompiled with extra defines, linked with the objects of the current build; measured by tools/variant_bench.py)
set -eu
name=$1; src=$2; shift 2
cd "$(dirname "$0")/../hip-bvh-construction_b200/csrc"
mkdir -p ../variants build/var
NVCC=/usr/local/cuda/bin/nvcc
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off "$@" -c $src -o build/var/$name.o
objs=$(ls *.o | grep -v "^${src%.cu}.o$")
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libb2bvh_$name.so $objs build/var/$name.o -lcudart
echo "built variants/libb2bvh_$name.so"
