#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python tests/perf/bench_extra.py > gpurun_out/bench_extra.json 2> gpurun_out/bench_extra.err
echo done
