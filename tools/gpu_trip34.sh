#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1
timeout 900 python tests/perf/bench_extra.py > gpurun_out/bench_extra.json 2> gpurun_out/bench_extra.err
echo done
