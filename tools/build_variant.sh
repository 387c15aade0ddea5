#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags]   -> tools/variants/libmtm_<name>.so
# Only estep.cu is recompiled ((4,4) kernels only); the other objects come from the in-tree build (csrc/_build).
set -e
name="$1"; shift
cd "$(dirname "$0")/../multimodal-trajectory-modeling_b200/csrc"
mkdir -p ../../tools/variants _build_var
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC $ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall -ccbin /usr/bin/g++ -DMTM_ONLY_44 "$@" -c estep.cu -o _build_var/estep_$name.o
$NVCC $ARCH -shared -o ../../tools/variants/libmtm_$name.so _build_var/estep_$name.o $(ls _build/*.o | grep -v "/estep.o") -ldl
echo built libmtm_$name.so
