#!/bin/bash
# One parameterised runner for everything that goes to the GPU box.  Usage, from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_trip.sh tests bench ncu_p5 ...'
# Each word is a stage; outputs land in gpurun_out/ (copy what is to be kept into profiles/r<NN>_*).  Stages are
# independent and each runs under its own `timeout`, so one hang cannot eat the lease.  A call costs ~3.5 min of
# budget before its first stage runs: batch stages.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
# captures look at ONE launch: the 4th update = search-heavy frame, the 130th = steady frame
BENCH_NCU="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
# tune_probe configuration = the library's defaults, spelt out (see tools/tune_probe.py)
DEF="16,512,384,16,128,50,1,8,1,0,0,2,64"
# A/B stages compare librmd_b200.so with variant builds: make them first, here, with tools/build_variant_lib.sh
#   nopair:  -DRMD_STAGED_PAIR_MAX_PS=0      strip40 / strip80:  -DRMD_STRIP_FLOATS=10240 / 20480
ab_libs() { for lib in "$@"; do echo "library $lib:"
              for cfg in 640,480,200,5 1280,720,300,5 1920,1080,120,7; do
                RMD_PROBE_SIZE=${cfg%,*} RMD_PROBE_PATCH=${cfg##*,} RMD_B200_LIB=$PWD/rpg_open_remode_b200/$lib timeout 600 python tools/tune_probe.py $DEF 2>&1 | tail -n 1 | sed "s/^.*wt_cands 64/  $cfg/"; done; done; }
for stage in "$@"; do
  echo "=== stage $stage $(date +%T)"
  case "$stage" in
    # ---- correctness
    tests)      timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log ;;
    tests_all)  timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log ;;
    parity)     timeout 1500 python -m pytest tests/test_full_length_parity.py -m gpu -q -s > $OUT/pytest_parity.log 2>&1; tail -5 $OUT/pytest_parity.log ;;
    cpp)        timeout 900 python -m pytest tests/test_cpp_facade.py tests/test_cpp_multi_gpu.py -m gpu -q > $OUT/pytest_cpp.log 2>&1; tail -5 $OUT/pytest_cpp.log ;;
    smoke)      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log ;;
    sanitize)   timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/batch_equality_probe.py diff 3 > $OUT/sanitize_memcheck.txt 2>&1; tail -n 15 $OUT/sanitize_memcheck.txt
                timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/batch_equality_probe.py same 2 > $OUT/sanitize_racecheck.txt 2>&1; tail -n 15 $OUT/sanitize_racecheck.txt ;;
    # ---- the bench lines (profiles/r02_bench_*.json)
    bench)      timeout 900 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; tail -c 600 $OUT/bench_c2.json ;;
    bench_c3)   timeout 900 python bench.py --config c3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err ;;
    bench_c4)   timeout 900 python bench.py --config c4 > $OUT/bench_c4.json 2> $OUT/bench_c4.err ;;
    bench_ref)  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_c2.json 2> $OUT/bench_ref_c2.err ;;
    bench_ref_c3) timeout 900 python bench.py --impl reference --config c3 --steps 2 --warmup 1 > $OUT/bench_ref_c3.json 2> $OUT/bench_ref_c3.err ;;
    bench_ref_c4) timeout 900 python bench.py --impl reference --config c4 --steps 1 --warmup 1 > $OUT/bench_ref_c4.json 2> $OUT/bench_ref_c4.err ;;
    extra)      timeout 900 python tests/perf/bench_extra.py > $OUT/bench_extra.json 2> $OUT/bench_extra.err ;;
    # ---- ncu: launch lists and full captures (summarise with tools/ncu_summary.py)
    launches)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches_c2.csv $BENCH_NCU > $OUT/launches_c2.log 2>&1 ;;
    launches_c3) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_c3.csv $BENCH_NCU --config c3 > $OUT/launches_c3.log 2>&1 ;;
    launches_ref) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_ref_c2.csv python bench.py --impl reference --steps 1 --warmup 0 > $OUT/launches_ref_c2.log 2>&1 ;;
    ncu_p5)     timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 3 -c 1 -f -o $OUT/prof_staged_p5_heavy $BENCH_NCU > $OUT/ncu_p5_heavy.log 2>&1
                timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 129 -c 1 -f -o $OUT/prof_staged_p5_steady $BENCH_NCU > $OUT/ncu_p5_steady.log 2>&1 ;;
    ncu_p7)     timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 3 -c 1 -f -o $OUT/prof_staged_p7_heavy $BENCH_NCU --config c4 --frames 140 > $OUT/ncu_p7_heavy.log 2>&1
                timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 129 -c 1 -f -o $OUT/prof_staged_p7_steady $BENCH_NCU --config c4 --frames 140 > $OUT/ncu_p7_steady.log 2>&1 ;;
    ncu_warm)   timeout 900 ncu --cache-control none --clock-control none --section WarpStateStats --section SchedulerStats -k regex:depth_filter_staged -s 129 -c 1 $BENCH_NCU 2>&1 | grep -v "^==PROF==" | tail -n 60 > $OUT/ncu_p5_steady_warm.txt; cat $OUT/ncu_p5_steady_warm.txt ;;
    ncu_denoise) timeout 900 ncu --set full --clock-control none --import-source on -k regex:denoise_ -s 2 -c 2 -f -o $OUT/prof_denoise python tests/perf/denoise_probe.py --once > $OUT/ncu_denoise.log 2>&1 ;;
    # ---- probes
    denoise)    timeout 600 python tests/perf/denoise_probe.py > $OUT/denoise_probe.txt 2>&1; tail -12 $OUT/denoise_probe.txt ;;
    timeline)   timeout 600 python tools/timeline_probe.py > $OUT/timeline.txt 2>&1 ;;
    multi_kf)   timeout 600 python tools/multi_keyframe_probe.py > $OUT/multi_keyframe_probe.txt 2>&1; cat $OUT/multi_keyframe_probe.txt ;;
    e2e_probe)  timeout 600 python tools/e2e_probe.py > $OUT/e2e_probe.txt 2>&1 ;;
    # ---- tuning sweeps and A/Bs (profiles/r02_tune_probe.txt, r02_occupancy_ab.txt)
    tune)       timeout 900 python tools/tune_probe.py > $OUT/tune_probe.txt 2>&1; cat $OUT/tune_probe.txt ;;
    occupancy_ab) for size in 640,480,200 1280,720,300; do echo "image, frames: $size"; RMD_PROBE_SIZE=$size timeout 600 python tools/tune_probe.py 2>&1 | tail -n 4; done ;;
    knobs)      # sparse-path threshold, splitting, heavy-list threshold, split target, warp tiles (seeds, candidates)
                B="16,512,384,16,128,50,1,8,1,0,0,2,64"
                timeout 900 python tools/tune_probe.py $B ${B/,16,128,50,/,32,128,50,} ${B/,16,128,50,/,64,128,50,} ${B/16,512,384,/16,256,128,} ${B/16,512,384,/16,384,256,} \
                  ${B/,128,50,/,32,50,} ${B/,128,50,/,64,50,} ${B/,128,50,/,256,50,} ${B/,128,50,/,128,100,} ${B/,128,50,/,128,200,} \
                  ${B/,1,8,1,0,0,2,64/,1,0,1,0,0,2,64} ${B/,1,8,1,0,0,2,64/,1,4,1,0,0,2,64} ${B/,1,8,1,0,0,2,64/,1,8,1,0,0,2,32} $B 2>&1 | tail -n 14 ;;
    seed_mode)  for pct in 0 15 25 40; do timeout 600 python bench.py --seed-mode-pct $pct --no-cpu-baseline --steps 10 > $OUT/bench_c2_seed$pct.json 2> $OUT/bench_c2_seed$pct.err
                  python -c "import json; d=json.loads(open('$OUT/bench_c2_seed$pct.json').read()); print('seed pct $pct: value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'], 2))"; done ;;
    chain_ab)   for n in 1 8; do timeout 900 python bench.py --config c4 --frames 120 --no-e2e --no-cpu-baseline --chain-frames $n > $OUT/bench_c4_chain$n.json 2> $OUT/bench_c4_chain$n.err
                  python -c "import json; print('c4, 120 frames, $n frame(s) per launch:', json.loads(open('$OUT/bench_c4_chain$n.json').read())['ms_per_step'], 'ms')"; done ;;
    pair_ab)    ab_libs librmd_b200.so build/librmd_b200_nopair.so ;;
    strip_ab)   ab_libs build/librmd_b200_strip40.so librmd_b200.so build/librmd_b200_strip80.so ;;
    *)          echo "unknown stage $stage" ;;
  esac
done
echo "=== done $(date +%T)"
