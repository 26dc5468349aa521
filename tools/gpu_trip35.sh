#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/ingest_probe.py > gpurun_out/ingest_probe.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"undistort|u8_to_float" -c 60 --csv --log-file gpurun_out/launches_ingest.csv python tools/ingest_probe.py 12 > gpurun_out/ingest_under_ncu.log 2>&1
echo done
