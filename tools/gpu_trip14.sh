#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/timeline_probe.py > gpurun_out/timeline.txt 2>&1
echo done
