#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tests/golden/make_golden.py > gpurun_out/make_golden.log 2>&1
cp tests/golden/ref_cuda_96x72.npz gpurun_out/ 2>/dev/null
timeout 1500 python -m pytest tests/test_gpu_configs.py -m gpu -q > gpurun_out/pytest_configs.log 2>&1
timeout 1500 python tests/perf/bench_extra.py > gpurun_out/bench_extra.json 2> gpurun_out/bench_extra.err
echo done
