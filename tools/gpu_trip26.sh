#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 39 -c 1 -f -o gpurun_out/prof_staged_f40 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_f40.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 129 -c 1 -f -o gpurun_out/prof_staged_f130 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_f130.log 2>&1
echo done
