#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 python tests/perf/bench_extra.py > gpurun_out/bench_extra.json 2> gpurun_out/bench_extra.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file gpurun_out/launches_staged.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 3 -c 1 -f -o gpurun_out/prof_staged_heavy python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_heavy.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 129 -c 1 -f -o gpurun_out/prof_staged_steady python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_steady.log 2>&1
timeout 600 python tools/timeline_probe.py > gpurun_out/timeline.txt 2>&1
timeout 600 python tools/e2e_probe.py > gpurun_out/e2e_probe.txt 2>&1
echo done
