#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
timeout 120 tools/build/copy_probe > gpurun_out/copy_probe.txt 2>&1
RMD_B200_LIB=$PWD/tools/build/librmd_b200_dbg.so timeout 600 python tools/timeline_probe.py > gpurun_out/timeline_dbg.txt 2>&1
echo done
