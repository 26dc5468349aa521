#!/bin/bash
# First GPU trip: texture probe, gpu tests (direct kernel), 3-way diagnostics, first bench lines, launch list.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
( nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tex_probe tools/tex_probe.cu && timeout 120 /tmp/tex_probe ) > gpurun_out/tex_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "not staged" > gpurun_out/pytest_gpu.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "not staged" >> gpurun_out/pytest_gpu_all.log 2>&1
timeout 900 python tests/perf/gpu_diag.py --width 320 --height 240 --frames 30 > gpurun_out/diag_qvga.json 2> gpurun_out/diag_qvga.err
timeout 900 python tests/perf/gpu_diag.py --width 320 --height 240 --frames 12 --frac-bits 0 > gpurun_out/diag_qvga_exact.json 2> gpurun_out/diag_qvga_exact.err
timeout 900 python tests/perf/gpu_diag.py --width 640 --height 480 --frames 40 --no-oracle > gpurun_out/diag_vga.json 2> gpurun_out/diag_vga.err
timeout 900 python bench.py --variant direct --steps 3 --warmup 3 > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_direct.csv python bench.py --variant direct --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo done
