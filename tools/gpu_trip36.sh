#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/ingest_probe.py > gpurun_out/ingest_probe.txt 2>&1
echo done
