#!/bin/bash
# Debug-counter build of the library (RMD_DEBUG_COUNTERS=1) for tools/timeline_probe.py:
#   RMD_B200_LIB=tools/build/librmd_b200_dbg.so python tools/timeline_probe.py
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/build/dbg
SRC=rpg_open_remode_b200/csrc
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -use_fast_math -Xcompiler -fPIC -DRMD_DEBUG_COUNTERS=1"
for f in c_api depth_filter depth_filter_staged denoiser reduction ingest point_cloud; do
  nvcc $FLAGS -c $SRC/$f.cu -o tools/build/dbg/$f.o &
done
wait
nvcc -shared -o tools/build/librmd_b200_dbg.so tools/build/dbg/*.o -gencode arch=compute_100a,code=sm_100a -lpthread
ls -la tools/build/librmd_b200_dbg.so
