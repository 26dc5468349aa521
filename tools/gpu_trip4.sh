#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/tma_probe.cu > gpurun_out/tma_probe2.txt 2>&1
for v in 0 7 8 9 1; do timeout 60 /tmp/tma_probe $v >> gpurun_out/tma_probe2.txt 2>&1; echo "exit $?" >> gpurun_out/tma_probe2.txt; done
nvidia-smi -q | grep -i -E "mig|virtual|confidential|persistence|compute mode" >> gpurun_out/tma_probe2.txt 2>&1
timeout 600 python -m pytest tests/test_cpp_facade.py -m gpu -q -x > gpurun_out/pytest_cpp.log 2>&1
echo done
