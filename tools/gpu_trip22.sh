#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python tools/tune_probe.py > gpurun_out/tune_probe.txt 2>&1
echo done
