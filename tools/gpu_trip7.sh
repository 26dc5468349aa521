#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/tma_probe.cu > gpurun_out/tma_probe5.txt 2>&1
for v in 14 15 16 17 18 19; do DUMP_DESC=1 timeout 60 /tmp/tma_probe $v >> gpurun_out/tma_probe5.txt 2>&1; echo "exit $?" >> gpurun_out/tma_probe5.txt; done
python - >> gpurun_out/tma_probe5.txt 2>&1 <<'PY'
import triton, os, glob
d=os.path.dirname(triton.__file__)
print(glob.glob(d+'/backends/nvidia/bin/*'))
import subprocess
for p in glob.glob(d+'/backends/nvidia/bin/ptxas*'):
    print(p, subprocess.run([p,'--version'],capture_output=True,text=True).stdout.strip().splitlines()[-1])
PY
echo done
