#!/bin/bash
# Builds another instance of libflowagg.so with extra -D flags into flow-pipeline_b200/_variants/<name>.so (git-ignored,
# travels to the GPU box); FLOWAGG_LIB selects it at run time.  Usage: profiles/build_variant.sh <name> [nvcc flags...]
set -e
cd "$(dirname "$0")/../flow-pipeline_b200/csrc"
name=$1; shift
mkdir -p ../_variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function \
    -Xptxas -v "$@" -shared -o ../_variants/$name.so flowagg.cu -ldl 2> ../_variants/$name.log
grep -A2 "AggConsumerILi1ELb0EEELi256" ../_variants/$name.log | grep -E "registers|spill" || true
