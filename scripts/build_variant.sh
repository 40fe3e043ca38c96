#!/bin/bash
# build a diagnostic / tuning variant of the library: scripts/build_variant.sh <suffix> [-DMACRO ...]
# select it at run time with BAJES_B200_LIB=bajes_b200/csrc/libbajes_b200_<suffix>.so
set -e
cd "$(dirname "$0")/../bajes_b200/csrc"
suf=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" -o libbajes_b200_${suf}.so capi.cu -ldl
echo built libbajes_b200_${suf}.so
