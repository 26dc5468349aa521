#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_node_logic.py -m gpu -q > gpurun_out/pytest_node.log 2>&1
timeout 60 python tools/multi_keyframe_probe.py > gpurun_out/multi_keyframe_probe.txt 2>&1
echo done
