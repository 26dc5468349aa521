#!/bin/bash
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tools/tma_probe.cu > gpurun_out/tma_probe6.txt 2>&1
for v in 20 21 22; do timeout 60 /tmp/tma_probe $v >> gpurun_out/tma_probe6.txt 2>&1; echo "exit $?" >> gpurun_out/tma_probe6.txt; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "staged" > gpurun_out/pytest_staged.log 2>&1
if grep -q "passed" gpurun_out/pytest_staged.log && ! grep -q "failed" gpurun_out/pytest_staged.log; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "first_update_parity and staged" > gpurun_out/sanitizer_staged.log 2>&1
  timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_cpp_facade.py -m gpu -q > gpurun_out/pytest_gpu_a.log 2>&1
  timeout 1500 python -m pytest tests/test_ref_cuda_parity.py -m gpu -q > gpurun_out/pytest_gpu_b.log 2>&1
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  timeout 900 python bench.py --variant staged --steps 5 --warmup 3 > gpurun_out/bench_staged.json 2> gpurun_out/bench_staged.err
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_staged.csv python bench.py --variant staged --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 5 -c 2 -o gpurun_out/prof_staged python bench.py --variant staged --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_staged.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 120 -c 2 -o gpurun_out/prof_staged_steady python bench.py --variant staged --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_staged_steady.log 2>&1
else
  timeout 1500 python -m pytest tests/test_cpp_facade.py -m gpu -q > gpurun_out/pytest_gpu_a.log 2>&1
fi
echo done
