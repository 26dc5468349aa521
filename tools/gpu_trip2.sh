#!/bin/bash
# Second GPU trip: staged kernel validation, full gpu test suite, bench (staged/direct/reference), ncu.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tex_probe tools/tex_probe.cu && timeout 120 /tmp/tex_probe ) > gpurun_out/tex_probe2.txt 2>&1
# staged kernel first, bounded, and once under memcheck on a tiny case
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "staged" > gpurun_out/pytest_staged.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "first_update_parity and staged" > gpurun_out/sanitizer_staged.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_all2.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --variant staged --steps 5 --warmup 3 > gpurun_out/bench_staged.json 2> gpurun_out/bench_staged.err
timeout 900 python bench.py --variant direct --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_direct2.json 2> gpurun_out/bench_direct2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference2.json 2> gpurun_out/bench_reference2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_staged.csv python bench.py --variant staged --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 5 -c 3 -o gpurun_out/prof_staged python bench.py --variant staged --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_staged.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 120 -c 2 -o gpurun_out/prof_staged_steady python bench.py --variant staged --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_staged_steady.log 2>&1
echo done
