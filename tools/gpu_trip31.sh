#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_full.log 2>&1
timeout 600 python tools/timeline_probe.py > gpurun_out/timeline.txt 2>&1
timeout 900 python tools/tune_probe.py > gpurun_out/tune_probe.txt 2>&1
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_staged.json 2> gpurun_out/bench_staged.err
echo done
