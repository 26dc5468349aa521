#!/bin/bash
# One parameterised runner for everything that goes to the GPU box (replaces the
# round-1 gpu_trip*.sh one-offs).  Usage, from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_trip.sh tests bench ncu_p5 ...'
# Each word is a stage; outputs land in gpurun_out/.  Stages are independent and
# each runs under its own `timeout`, so one hang cannot eat the lease.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out
# captures look at ONE frame per launch (--chain-frames 1): the 4th update = search-heavy, the 130th = steady;
# stage ncu_chain captures a chained launch (8 frames) of the timed path as well
BENCH_NCU="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --chain-frames 1"
for stage in "$@"; do
  echo "=== stage $stage $(date +%T)"
  case "$stage" in
    tests)      timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log ;;
    tests_all)  timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log ;;
    parity)     timeout 1500 python -m pytest tests/test_full_length_parity.py -m gpu -q -s > $OUT/pytest_parity.log 2>&1; tail -5 $OUT/pytest_parity.log ;;
    smoke)      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log ;;
    bench)      timeout 900 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; tail -c 600 $OUT/bench_c2.json ;;
    bench_seedpct) for pct in 0 15 25 40; do timeout 600 python bench.py --seed-mode-pct $pct --no-cpu-baseline --steps 10 > $OUT/bench_c2_seed$pct.json 2> $OUT/bench_c2_seed$pct.err; python -c "import json; d=json.loads(open('gpurun_out/bench_c2_seed$pct.json').read()); print('seed pct $pct: value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],2), 'fused launches/step', d['gpu_launches_fused']/10)"; done ;;
    pair_ab)    for lib in librmd_b200.so build/librmd_b200_nopair.so; do echo "library $lib:"
                  for cfg in 640,480,200,5 640,480,200,7 1280,720,300,5 1920,1080,120,7; do
                    RMD_PROBE_SIZE=${cfg%,*} RMD_PROBE_PATCH=${cfg##*,} RMD_B200_LIB=$PWD/rpg_open_remode_b200/$lib timeout 600 python tools/tune_probe.py 16,512,384,16,32,100,1,8,1,0,0,2 2>&1 | tail -n 1 | sed "s/^.*ctas.sm 2/  $cfg/"; done; done ;;
    knobs)      T=",1,8,1,0,0,2"; timeout 600 python tools/tune_probe.py 16,512,384,16,32,100$T 16,512,384,24,32,100$T 16,512,384,32,32,100$T 16,512,384,48,32,100$T 16,512,384,64,32,100$T \
                  16,256,128,16,32,100$T 32,256,128,16,32,100$T 16,384,256,16,32,100$T 16,512,384,16,32,50$T 16,512,384,16,32,200$T 16,512,384,16,16,100$T 16,512,384,16,64,100$T 16,512,384,16,32,100$T 2>&1 | tail -n 13 | sed 's/min_items/mi/;s/per_cta/pc/;s/heavy_min/hm/;s/avg_pct/ap/;s/pdl 1 warp_tiles 8 chain 1 seed_pct 0 grid 0 ctas.sm 2//' ;;
    knobs2)     T=",1,8,1,0,0,2"; timeout 600 python tools/tune_probe.py 16,512,384,16,32,100$T 16,512,384,16,64,100$T 16,512,384,16,96,100$T 16,512,384,16,128,100$T 16,512,384,16,192,100$T 16,512,384,16,256,100$T 16,512,384,16,512,100$T \
                  16,512,384,16,64,50$T 16,512,384,16,128,50$T 16,512,384,16,64,70$T 16,384,256,16,64,100$T 16,384,256,16,128,70$T 16,512,384,16,32,100$T 2>&1 | tail -n 13 | sed 's/min_items/mi/;s/per_cta/pc/;s/heavy_min/hm/;s/avg_pct/ap/;s/pdl 1 warp_tiles 8 chain 1 seed_pct 0 grid 0 ctas.sm 2//' ;;
    rolled_ab)  T=",1,8,1,0,0,2"; for lib in librmd_b200.so build/librmd_b200_rolled.so; do echo "library $lib:"
                  RMD_B200_LIB=$PWD/rpg_open_remode_b200/$lib timeout 600 python tools/tune_probe.py 16,512,384,16,32,100$T 16,512,384,16,64,100$T 2>&1 | tail -n 2 | sed 's/^.*ctas.sm 2/  VGA/'
                  RMD_PROBE_SIZE=1280,720,300 RMD_B200_LIB=$PWD/rpg_open_remode_b200/$lib timeout 600 python tools/tune_probe.py 16,512,384,16,32,100$T 16,512,384,16,64,100$T 2>&1 | tail -n 2 | sed 's/^.*ctas.sm 2/  720p/'; done ;;
    ncu_warm)   timeout 900 ncu --cache-control none --clock-control none --section WarpStateStats --section SchedulerStats -k regex:depth_filter_staged -s 129 -c 1 $BENCH_NCU 2>&1 | grep -v "^==PROF==" | tail -n 60 > $OUT/ncu_p5_steady_warm.txt; cat $OUT/ncu_p5_steady_warm.txt ;;
    knobs3)     H="16,512,384,16,128,50,1"; T=",1,0,0,2"; for size in 640,480,200 1280,720,300; do echo "image, frames: $size"
                  RMD_PROBE_SIZE=$size timeout 600 python tools/tune_probe.py $H,8$T,64 $H,8$T,128 $H,12$T,96 $H,16$T,128 $H,16$T,256 $H,24$T,192 $H,32$T,256 $H,32$T,512 $H,0$T,64 $H,8$T,64 2>&1 | tail -n 10 | sed 's/^.*pdl 1 warp_tiles/  warp_tiles/;s/chain 1 seed_pct 0 grid 0 ctas.sm 2 //'; done ;;
    strip_ab)   H="16,512,384,16,128,50,1,8,1,0,0,2,64"; for lib in librmd_b200.so build/librmd_b200_strip60.so build/librmd_b200_strip80.so; do echo "library $lib:"
                  for cfg in 640,480,200,5 1280,720,300,5 1920,1080,120,7; do
                    RMD_PROBE_SIZE=${cfg%,*} RMD_PROBE_PATCH=${cfg##*,} RMD_B200_LIB=$PWD/rpg_open_remode_b200/$lib timeout 600 python tools/tune_probe.py $H 2>&1 | tail -n 1 | sed "s/^.*wt_cands 64/  $cfg/"; done; done ;;
    occupancy_ab) for size in 640,480,200 1280,720,300; do echo "image, frames: $size"; RMD_PROBE_SIZE=$size timeout 600 python tools/tune_probe.py 2>&1 | tail -n 4; done ;;
    bench_c3)   timeout 900 python bench.py --config c3 > $OUT/bench_c3.json 2> $OUT/bench_c3.err ;;
    bench_c4)   timeout 900 python bench.py --config c4 > $OUT/bench_c4.json 2> $OUT/bench_c4.err ;;
    bench_c4_ab) timeout 900 python bench.py --config c4 --frames 120 --no-e2e --no-cpu-baseline --chain-frames 1 > $OUT/bench_c4_chain1.json 2> $OUT/bench_c4_chain1.err
                timeout 900 python bench.py --config c4 --frames 120 --no-e2e --no-cpu-baseline --chain-frames 8 > $OUT/bench_c4_chain8.json 2> $OUT/bench_c4_chain8.err
                python -c "import json; [print(n, json.loads(open('gpurun_out/bench_c4_chain%d.json' % n).read())['ms_per_step']) for n in (1, 8)]" ;;
    bench_ref)  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_c2.json 2> $OUT/bench_ref_c2.err ;;
    bench_ref_c3) timeout 900 python bench.py --impl reference --config c3 --steps 2 --warmup 1 > $OUT/bench_ref_c3.json 2> $OUT/bench_ref_c3.err ;;
    bench_ref_c4) timeout 900 python bench.py --impl reference --config c4 --steps 1 --warmup 1 > $OUT/bench_ref_c4.json 2> $OUT/bench_ref_c4.err ;;
    launches)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/launches_c2.csv $BENCH_NCU > $OUT/launches_c2.log 2>&1 ;;
    launches_ref) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_ref_c2.csv python bench.py --impl reference --steps 1 --warmup 0 > $OUT/launches_ref_c2.log 2>&1 ;;
    launches_c3) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_c3.csv $BENCH_NCU --config c3 > $OUT/launches_c3.log 2>&1 ;;
    ncu_p5)     timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 3 -c 1 -f -o $OUT/prof_staged_p5_heavy $BENCH_NCU > $OUT/ncu_p5_heavy.log 2>&1
                timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 129 -c 1 -f -o $OUT/prof_staged_p5_steady $BENCH_NCU > $OUT/ncu_p5_steady.log 2>&1 ;;
    ncu_p7)     timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 3 -c 1 -f -o $OUT/prof_staged_p7_heavy $BENCH_NCU --config c4 --frames 140 > $OUT/ncu_p7_heavy.log 2>&1
                timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 129 -c 1 -f -o $OUT/prof_staged_p7_steady $BENCH_NCU --config c4 --frames 140 > $OUT/ncu_p7_steady.log 2>&1 ;;
    ncu_chain)  timeout 900 ncu --set full --clock-control none --import-source on -k regex:depth_filter_staged -s 16 -c 1 -f -o $OUT/prof_staged_p5_chain python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_p5_chain.log 2>&1 ;;
    ncu_denoise) timeout 900 ncu --set full --clock-control none --import-source on -k regex:denoise_ -s 2 -c 2 -f -o $OUT/prof_denoise python tools/denoise_probe.py --once > $OUT/ncu_denoise.log 2>&1 ;;
    denoise)    timeout 600 python tools/denoise_probe.py > $OUT/denoise_probe.txt 2>&1; tail -12 $OUT/denoise_probe.txt ;;
    timeline)   timeout 600 python tools/timeline_probe.py > $OUT/timeline.txt 2>&1 ;;
    tune)       timeout 900 python tools/tune_probe.py > $OUT/tune_probe.txt 2>&1; cat $OUT/tune_probe.txt ;;
    multi_kf)   timeout 600 python tools/multi_keyframe_probe.py > $OUT/multi_keyframe_probe.txt 2>&1; cat $OUT/multi_keyframe_probe.txt ;;
    extra)      timeout 900 python tests/perf/bench_extra.py > $OUT/bench_extra.json 2> $OUT/bench_extra.err ;;
    parity_diag) timeout 900 python tools/parity_diag.py > $OUT/parity_diag.txt 2>&1; tail -n 30 $OUT/parity_diag.txt ;;
    batch_debug) timeout 300 python tools/batch_debug.py same 2 > $OUT/batch_debug.txt 2>&1
                timeout 300 python tools/batch_debug.py diff 3 >> $OUT/batch_debug.txt 2>&1
                RMD_FORCE_BATCH_KERNEL=1 timeout 600 python -m pytest tests/test_gpu_worklist.py -m gpu -q -x >> $OUT/batch_debug.txt 2>&1
                tail -n 40 $OUT/batch_debug.txt ;;
    chain)      timeout 600 python -m pytest tests/test_full_length_parity.py -m gpu -q -x -k device_resident > $OUT/pytest_chain.log 2>&1; tail -n 15 $OUT/pytest_chain.log ;;
    batch_debug2) timeout 300 python tools/batch_debug2.py > $OUT/batch_debug2.txt 2>&1; cat $OUT/batch_debug2.txt ;;
    sanitize)   timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/batch_debug.py diff 3 > $OUT/sanitize_memcheck.txt 2>&1; tail -n 15 $OUT/sanitize_memcheck.txt
                timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/batch_debug.py same 2 > $OUT/sanitize_racecheck.txt 2>&1; tail -n 15 $OUT/sanitize_racecheck.txt ;;
    e2e_probe)  timeout 600 python tools/e2e_probe.py > $OUT/e2e_probe.txt 2>&1 ;;
    cpp)        timeout 900 python -m pytest tests/test_cpp_facade.py tests/test_cpp_multi_gpu.py -m gpu -q > $OUT/pytest_cpp.log 2>&1; tail -5 $OUT/pytest_cpp.log ;;
    *)          echo "unknown stage $stage" ;;
  esac
done
echo "=== done $(date +%T)"
