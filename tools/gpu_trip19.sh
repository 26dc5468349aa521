#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_full.log 2>&1
timeout 600 python tools/timeline_probe.py > gpurun_out/timeline.txt 2>&1
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_staged.json 2> gpurun_out/bench_staged.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_staged.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
echo done
