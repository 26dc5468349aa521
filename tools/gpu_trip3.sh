#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/tma_probe.cu > gpurun_out/tma_probe.txt 2>&1
for v in 1 2 3 4 5 6; do timeout 60 /tmp/tma_probe $v >> gpurun_out/tma_probe.txt 2>&1; echo "exit $?" >> gpurun_out/tma_probe.txt; done
timeout 1500 python -m pytest tests -m gpu -q -k "not staged" > gpurun_out/pytest_gpu_all3.log 2>&1
echo done
