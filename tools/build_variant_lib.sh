#!/bin/bash
# Builds a variant of librmd_b200.so with extra -D flags for the staged kernel (A/B experiments on the GPU box):
#   tools/build_variant_lib.sh mb2 -DRMD_STAGED_P5_MIN_BLOCKS=2
# -> rpg_open_remode_b200/build/librmd_b200_<tag>.so; load it with RMD_B200_LIB=<path> (rpg_open_remode_b200/_native.py).
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
python -c "from rpg_open_remode_b200 import _build; _build.build_cuda()"
B=rpg_open_remode_b200/build
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -use_fast_math -Xcompiler -fPIC "$@" \
     -c rpg_open_remode_b200/csrc/depth_filter_staged.cu -o $B/depth_filter_staged_$tag.o
objs=$(ls $B/*.o | grep -v "depth_filter_staged" | grep -v "_$tag.o" | grep -v "staged_")
nvcc -shared -o $B/librmd_b200_$tag.so $objs $B/depth_filter_staged_$tag.o -gencode arch=compute_100a,code=sm_100a -lpthread -ldl
echo "$B/librmd_b200_$tag.so"
