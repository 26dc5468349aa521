#!/usr/bin/env python
"""bench.py -- depth-filter frames/sec on B200 (BASELINE.json metric).

Workloads (--config, default c2 = the configuration BASELINE.json's metric is quoted on):

  c2  640x480 synthetic pinhole sequence, 200 frames, 5x5 NCC            (BASELINE configs[1])
  c3  1280x720, 500 frames, 5x5 NCC + TV-L1 denoiser, 50 iterations       (configs[2])
  c4  1920x1080, 500 frames, 7x7 NCC                                      (configs[3])
  c5  = c2 run with --gpus N: N independent keyframes, one per GPU        (configs[4])

A *step* is one pass of the hot path over one such sequence: setReferenceImage on
frame 0, then one fused depth-filter update (check + epipolar NCC search +
triangulation + Bayesian update) per remaining frame, then (c3) the denoiser, then
the final gather of the depth and convergence maps to rank 0.  frames/sec counts
the update() frames, the protocol of the reference's test/dataset_main.cpp:101-105.

  value   whole-job frames/s with the frames already resident in HBM (float frames,
          > the 126 MB L2 per step, streamed once; the seed state is legitimately
          L2-resident from one frame to the next -- that is the workload);
  e2e     the same metric through the reference-facing host API with HOST buffers:
          rmd::Depthmap's protocol (src/depthmap.cpp:63-93) -- 8-bit gray frames in
          page-locked host memory -> setReferenceImage / update -> H2D copy and the
          8U->32F conversion inside the timed region -> final depth + convergence
          maps read back to the host every step;
  N > 1   independent keyframes, one per rank (weak scaling), NCCL only for the
          final gather of the depth and convergence maps to rank 0.

`--impl reference` runs the reference's own implementation of the path instead:
its unmodified CUDA kernels rebuilt for sm_100a (oracle/_ref/librmd_ref*.so,
recipe oracle/Makefile) through rmd::SeedMatrix with host buffers (8-bit frames
converted on the host like rmd::Depthmap::inputImage, src/depthmap.cpp:105), when
that library and a GPU are present; otherwise the CPU oracle port.  (The reference
has no CPU implementation of the path; its "dataset_main" drives CUDA kernels:
SURVEY.md, first table.)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

UNIT = "frames/s"
BYTES_PER_PIXEL_FUSED = 52        # algorithmic bytes per pixel per frame of the fused kernel
BYTES_PER_PIXEL_DENOISE_IT = 40   # per pixel per primal-dual iteration

CONFIGS = {
    # name: width, height, frames, patch, denoise iterations, synthetic seed (None: per-keyframe), steps, label
    "c2": dict(width=640, height=480, frames=200, patch=5, denoise=0, seed=None, steps=20,
               label="BASELINE configs[1]", metric="depth-filter frames/sec (VGA, 200-frame seq)"),
    "c3": dict(width=1280, height=720, frames=500, patch=5, denoise=50, seed=0x5EED0003, steps=4,
               label="BASELINE configs[2]", metric="depth-filter frames/sec (720p, 500-frame seq + TV-L1 50 it)"),
    "c4": dict(width=1920, height=1080, frames=500, patch=7, denoise=0, seed=0x5EED0004, steps=3,
               label="BASELINE configs[3]", metric="depth-filter frames/sec (1080p, 500-frame seq, 7x7 NCC)"),
    "c5": dict(width=640, height=480, frames=200, patch=5, denoise=0, seed=None, steps=20,
               label="BASELINE configs[4] (= configs[1] per GPU; run with --gpus 8)",
               metric="depth-filter frames/sec (VGA, 200-frame seq)"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--patch", type=int, default=None)
    ap.add_argument("--denoise-iters", type=int, default=None)
    ap.add_argument("--variant", default=None, choices=[None, "staged", "direct"])
    ap.add_argument("--chain-frames", type=int, default=None,
                    help="frames per chained launch of the device-resident path (RMD_OPT_CHAIN_FRAMES, 1..8; default: the library's)")
    ap.add_argument("--seed-mode-pct", type=int, default=None,
                    help="RMD_OPT_SEED_MODE_PCT: go seed-major when at most this percentage of the pixels is still updated")
    ap.add_argument("--cpu-frames", type=int, default=30,
                    help="bounded sample of the sequence for the CPU baseline (frames 1..n)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the end-to-end and extra passes")
    ap.add_argument("--reference-cpu", action="store_true",
                    help="--impl reference: force the CPU oracle port even if the reference CUDA build exists")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    for k in ("width", "height", "frames", "patch"):
        if getattr(args, k) is None:
            setattr(args, k, cfg[k])
    if args.denoise_iters is None:
        args.denoise_iters = cfg["denoise"]
    if args.steps is None:
        args.steps = cfg["steps"]
    args.metric = cfg["metric"]
    args.label = cfg["label"]
    args.seed = cfg["seed"]
    return args


# ------------------------------------------------------------------ helpers

class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def sequence_seed(args, rank):
    from rpg_open_remode_b200 import multi_gpu
    # one keyframe per rank: rank r owns keyframe r (multi_gpu.shard_keyframes(world, r, world) == [r])
    return multi_gpu.keyframe_seed(rank) if args.seed is None else args.seed + 16 * rank


def load_sequence(args, rank, want_float=True):
    """Renders the rank's synthetic sequence on the host: 8-bit frames (the reference's MONO8 input) and,
    optionally, the float frames u8 * (1/255.f) of src/depthmap.cpp:105."""
    from rpg_open_remode_b200 import synth
    seq = synth.SyntheticSequence(args.width, args.height, seed=sequence_seed(args, rank))
    n = args.frames
    frames_u8 = np.empty((n, args.height, args.width), np.uint8)
    frames = np.empty((n, args.height, args.width), np.float32) if want_float else None
    poses = np.empty((n, 12), np.float32)   # T_curr_world (world -> camera)
    depth0 = None
    for k in range(n):
        f = seq.frame(k, want_depth=(k == 0))
        frames_u8[k] = f.image_u8
        if want_float:
            frames[k] = f.image
        poses[k] = f.T_cam_world.reshape(12)
        if k == 0:
            depth0 = f.depth
    return seq, frames_u8, frames, poses, float(depth0.min()), float(depth0.max())


def _profile(name):
    """Newest committed profile of that name (profiles/r02_* before r01_*)."""
    for rnd in ("r02", "r01"):
        path = os.path.join(ROOT, "profiles", f"{rnd}_{name}")
        if os.path.exists(path):
            return path
    return None


def ncu_kernel_share(patch):
    """Share of the step's kernel time spent in the fused depth-filter kernel, from the committed ncu
    launch list of this same command; None if absent."""
    import csv
    path = _profile("launches_staged_kernel.csv") if patch == 5 else _profile("launches_staged_kernel_p7.csv")
    if not path:
        return None
    try:
        with open(path) as f:
            rows = list(csv.reader(f))
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        hdr = rows[hi]
        kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3}
        fused = total = 0.0
        for r in rows[hi + 1:]:
            if len(r) <= mv:
                continue
            v = float(r[mv].replace(",", "")) * scale.get(r[mu], 1.0)
            total += v
            if "depth_filter_" in r[kn]:
                fused += v
        return {"value": fused / total, "source": os.path.relpath(path, ROOT)} if total > 0 else None
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_capture(patch, algo_bytes=None):
    """The committed `ncu --set full` capture of the dominant kernel (heavy and steady frame): DRAM traffic
    per launch and the issue-slot utilisation that actually bounds it.  Never measured under the timed run.
    Captures exist for the VGA 5x5 and the 1080p 7x7 workloads: for any other image size the traffic is not
    reported (None) rather than borrowed from a capture of different launches."""
    path = _profile("ncu_staged.json") if patch == 5 else _profile("ncu_staged_p7.json")
    if not path:
        return None, None
    try:
        with open(path) as f:
            j = json.load(f)
        if algo_bytes is not None and abs(float(j.get("algorithmic_bytes_per_launch", 0.0)) - float(algo_bytes)) > 0.5:
            return None, None
        t = j["traffic_bytes_per_launch"]
        src = os.path.relpath(path, ROOT)
        traffic = {"heavy_frame": t["heavy"], "steady_frame": t["steady"], "source": src}
        m = j["metrics"]
        key = "smsp__issue_active.avg.pct_of_peak_sustained_active"
        dur = m["gpu__time_duration.sum"]
        to_us = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6}
        bound = {"bound": "dependency latency at 16 warps per SM (search-heavy frames) / critical path of the busiest "
                          "CTA (steady frames)",
                 "heavy_frame_issue_slots_pct": float(m[key]["heavy"]),
                 "steady_frame_issue_slots_pct": float(m[key]["steady"]),
                 "heavy_frame_us_under_ncu": float(dur["heavy"]) * to_us.get(dur.get("unit", "us"), 1.0),
                 "steady_frame_us_under_ncu": float(dur["steady"]) * to_us.get(dur.get("unit_steady", dur.get("unit", "us")), 1.0),
                 "source": src}
        return traffic, bound
    except Exception:
        return None, None


def cpu_baseline(args, frames, poses, dmin, dmax, seq):
    """CPU oracle port (oracle/librmd_oracle.so, OpenMP, all host threads) timed on a bounded sample of the
    same workload: frames 1..n of the sequence, after one untimed warm-up pass over the first frames (thread
    pool start-up and page faults made the round-1 number vary 3x between runs)."""
    import oracle_binding as ob
    n = max(1, min(args.cpu_frames, args.frames - 1))
    threads = ob.get_threads()
    o = ob.OracleSeeds(args.width, args.height, *seq.camera, patch=args.patch)
    o.set_reference(frames[0], poses[0], dmin, dmax)
    for k in range(1, min(3, n) + 1):
        o.update(frames[k], poses[k])
    o.set_reference(frames[0], poses[0], dmin, dmax)
    t0 = time.perf_counter()
    for k in range(1, n + 1):
        o.update(frames[k], poses[k])
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"frames 1..{n} of the {args.width}x{args.height} sequence (the search-heavy start), "
                      f"{dt:.1f} s wall after a 3-frame warm-up, OpenMP over image rows"}


def workload_text(args, n_upd, per_gpu=True):
    s = (f"{args.width}x{args.height} synthetic pinhole sequence, {args.frames} frames ({n_upd} updates), "
         f"{args.patch}x{args.patch} NCC")
    if args.denoise_iters:
        s += f", TV-L1 denoiser {args.denoise_iters} iterations (lambda 0.5)"
    s += ", 1 reference keyframe" + (" per GPU" if per_gpu else "") + f" ({args.label})"
    return s


# ------------------------------------------------------------------- ours

def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import rpg_open_remode_b200 as rmd
    from rpg_open_remode_b200 import multi_gpu

    if not torch.cuda.is_available() or rmd.device_count() < 1:
        raise RuntimeError("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    seq, frames_u8, frames, poses, dmin, dmax = load_sequence(args, rank)
    W, H, NF = args.width, args.height, args.frames
    n_upd = NF - 1
    iters = args.denoise_iters

    # rmd::Depthmap's shape: one SeedMatrix + one DepthmapDenoiser (include/rmd/depthmap.h:107,123)
    fx, fy, cx, cy = seq.camera
    dm = rmd.Depthmap(W, H, fx, cx, fy, cy, patch_side=args.patch, device=local_rank)
    seeds, den = dm.seeds_, dm.denoiser_
    variant = args.variant or os.environ.get("RMD_BENCH_VARIANT", "staged")
    seeds.setOption(rmd.OPT_KERNEL_VARIANT, rmd.VARIANT_STAGED if variant == "staged" else rmd.VARIANT_DIRECT)
    # e2e: frames live in page-locked host memory and are DMA'd in place (the contract's "host->device copy of
    # that step's inputs from pinned host memory"); pageable callers still get the staged, reusable-on-return path
    seeds.setOption(rmd.OPT_PINNED_INPUT, 1)
    if args.chain_frames:
        seeds.setOption(rmd.OPT_CHAIN_FRAMES, args.chain_frames)
    if args.seed_mode_pct is not None:
        seeds.setOption(rmd.OPT_SEED_MODE_PCT, args.seed_mode_pct)
    # A dedicated (non-default) torch stream is made current and handed to the handles, so the fused
    # kernels, the denoiser, torch's NCCL calls and the CUDA events that time them are all on one stream.
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    seeds.setStream(stream.cuda_stream)
    den.setStream(stream.cuda_stream)
    den.setLargeSigmaSq(dmax - dmin)

    # float frames resident in HBM for `value`; page-locked 8-bit host frames for `e2e`
    dev_frames = torch.from_numpy(frames).to(dev, non_blocking=False)
    pinned_u8 = torch.from_numpy(frames_u8).pin_memory()
    host_u8 = [pinned_u8[k].numpy() for k in range(NF)]
    frame_bytes = W * H * 4
    gatherer = multi_gpu.MapGatherer(H, W, dev, dst=0)    # depth + convergence in one buffer: one collective per step
    depth_out, conv_out = gatherer.depth, gatherer.convergence
    host_maps = torch.empty((2, H, W), dtype=torch.int32).pin_memory()

    def final_maps():
        # final depth map (denoised for c3) and convergence map into the gather buffer
        if iters:
            den.denoiseSeedsToDevice(seeds, depth_out.data_ptr(), W * 4, 0.5, iters)
        else:
            seeds.copyFieldToDevice(rmd.FIELD_MU, depth_out.data_ptr(), W * 4)
        seeds.copyFieldToDevice(rmd.FIELD_CONVERGENCE, conv_out.data_ptr(), W * 4)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step_resident(marks=None):
        if marks is not None:
            marks[0].record(stream)
        seeds.setReferenceImageDevice(dev_frames[0].data_ptr(), W * 4, poses[0], dmin, dmax)
        if marks is not None:
            marks[1].record(stream)
        seeds.updateDeviceBatch(dev_frames[1].data_ptr(), frame_bytes, W * 4, poses[1:])
        if marks is not None:
            marks[2].record(stream)
        final_maps()
        if marks is not None:
            marks[3].record(stream)
        gatherer.gather()   # the only collective on the path: final maps to rank 0 (NCCL)
        if marks is not None:
            marks[4].record(stream)

    def step_e2e():
        dm.denoiser_.setLargeSigmaSq(dmax - dmin)
        seeds.setReferenceImage(host_u8[0], poses[0], dmin, dmax)
        for k in range(1, NF):
            seeds.update(host_u8[k], poses[k])
        final_maps()
        gatherer.gather()
        # device -> host read of the step's result (final maps of this rank)
        host_maps.copy_(gatherer.packed, non_blocking=True)
        stream.synchronize()
        return host_maps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        return multi_gpu.max_over_ranks(v, dev)

    # ---------------- the two timed paths must produce the same maps (checked once, before timing)
    step_resident()
    torch.cuda.synchronize(dev)
    resident_maps = gatherer.packed.cpu().numpy().copy()
    if not args.no_e2e:
        e2e_maps = step_e2e().numpy()
        if not np.array_equal(resident_maps, e2e_maps):
            diff = int((resident_maps != e2e_maps).sum())
            raise RuntimeError(f"bench.py: the device-resident step and the end-to-end step disagree in {diff} map "
                               "elements -- refusing to time paths that do not compute the same thing")

    # ---------------- device-resident timing (value)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0, dl0 = seeds.launchCount(), den.launchCount()
    marks = [[ev() for _ in range(5)] for _ in range(args.steps)]
    e0, e1 = ev(), ev()
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step_resident(marks[i])
    e1.record(stream)
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches1, dl1 = seeds.launchCount(), den.launchCount()
    clocks = sampler.stop() if rank == 0 else None
    value = world * n_upd * args.steps / (dev_ms * 1e-3)
    ms_per_step = dev_ms / args.steps
    seg = np.array([[m[i].elapsed_time(m[i + 1]) for i in range(4)] for m in marks]).mean(axis=0)
    breakdown = {"set_reference_ms": float(seg[0]), "updates_ms": float(seg[1]),
                 "final_maps_ms": float(seg[2]), "gather_ms": float(seg[3])}
    if world > 1:
        t = torch.tensor(list(breakdown.values()), dtype=torch.float64, device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = np.stack([x.cpu().numpy() for x in allt])
        breakdown = {k: [float(v) for v in per_rank[:, i]] for i, k in enumerate(breakdown)}
        breakdown["updates_ms_max_minus_min"] = float(per_rank[:, 1].max() - per_rank[:, 1].min())
        breakdown["note"] = ("per rank; rank r runs keyframe r (different synthetic scene seeds), so updates_ms differs "
                             "per rank and the step is paced by the slowest keyframe: rank 0's gather waits for it")

    # ---------------- per-launch duration of the dominant kernel (roofline)
    # `avg_launch_us`: device time of the update segment of the timed steps (events around updateDeviceBatch
    # only: no seed-init, no export, no gather) / fused launches in it.  The event-pair pass below is the
    # distribution; a pair around every launch suppresses the launch overlap (PDL), so its mean is larger.
    seeds.setReferenceImageDevice(dev_frames[0].data_ptr(), W * 4, poses[0], dmin, dmax)
    evs = [(ev(), ev()) for _ in range(n_upd)]
    for k in range(1, NF):
        a, b = evs[k - 1]
        a.record(stream)
        seeds.updateDevice(dev_frames[k].data_ptr(), W * 4, poses[k])
        b.record(stream)
    torch.cuda.synchronize(dev)
    per_launch_ms = np.array([a.elapsed_time(b) for a, b in evs])
    conv_hist = np.bincount(seeds.downloadConvergence().ravel(), minlength=6).tolist()
    avg_launch_s = float(seg[1]) * 1e-3 / n_upd
    algo_bytes = BYTES_PER_PIXEL_FUSED * W * H
    peak, peak_src = peaks()
    achieved = algo_bytes / avg_launch_s / 1e9
    traffic, bound = ncu_capture(args.patch, algo_bytes)
    roofline = {"bound": "hbm", "kernel": "depth_filter_%s_kernel<%d>" % (variant, args.patch),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": (traffic or {}).get("heavy_frame"),
                "traffic_detail": traffic,
                "algorithmic_bytes_per_launch": algo_bytes,
                "avg_launch_us": avg_launch_s * 1e6,
                "avg_launch_us_definition": "CUDA events around the update segment of every timed step / fused launches",
                "step_derived_us": ms_per_step * 1e3 / n_upd,
                "per_launch_event_pass": {
                    "avg_us": float(per_launch_ms.mean() * 1e3),
                    "min_median_max_us": [float(per_launch_ms.min() * 1e3), float(np.median(per_launch_ms) * 1e3),
                                          float(per_launch_ms.max() * 1e3)],
                    "first_19_frames_ms": float(per_launch_ms[:19].sum()),
                    "rest_ms": float(per_launch_ms[19:].sum())},
                "kernel_share_of_step": ncu_kernel_share(args.patch),
                "what_bounds_it": bound,
                "note": "search-heavy frames are bound by dependency latency (<=143 candidates x P^2 bilinear taps per "
                        "seed, sums serial by bit-parity), steady frames by the critical path of the busiest CTA; not HBM "
                        "bound: see DESIGN.md 4.1 'What bounds it'"}

    out = {
        "metric": args.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args, n_upd),
                   "name": args.config,
                   "parallelism": f"{world} independent keyframes, NCCL gather of final depth+convergence",
                   "kernel_variant": variant,
                   "frames_per_launch_resident": args.chain_frames or 1,
                   "l2": "inputs larger than L2: %d distinct frames = %.0f MB streamed per step; seed state "
                         "(%.1f MB) is L2-resident by design" % (NF, NF * frame_bytes / 1e6, 28 * W * H / 1e6),
                   "final_state_hist[update,converged,border,diverged,no_match,not_visible]": conv_hist},
        "roofline": roofline,
        "breakdown_ms_per_step": breakdown,
        "gpu_launches": int((launches1[1] - launches0[1]) + (dl1 - dl0)),
        "gpu_launches_fused": int((launches1[0] - launches0[0])),
        "clocks": clocks,
    }

    # ---------------- the denoiser alone (north-star kernel): device time, 40 B per pixel per iteration
    if not args.no_e2e or iters:
        dn = {}
        for it in sorted({50, 200} | ({iters} if iters else set())):
            for _ in range(2):
                den.denoiseSeedsToDevice(seeds, depth_out.data_ptr(), W * 4, 0.5, it)
            a, b = ev(), ev()
            reps = 5
            a.record(stream)
            for _ in range(reps):
                den.denoiseSeedsToDevice(seeds, depth_out.data_ptr(), W * 4, 0.5, it)
            b.record(stream)
            torch.cuda.synchronize(dev)
            ms = a.elapsed_time(b) / reps
            gbs = BYTES_PER_PIXEL_DENOISE_IT * W * H * it / (ms * 1e-3) / 1e9
            dn[str(it)] = {"ms": ms, "us_per_iteration": ms * 1e3 / it, "achieved": gbs, "frac": gbs / peak}
        out["roofline_denoiser"] = {"bound": "hbm", "kernel": "denoise kernels (setup + iterations + export)",
                                    "unit": "GB/s", "peak": peak,
                                    "algorithmic_bytes_per_pixel_per_iteration": BYTES_PER_PIXEL_DENOISE_IT,
                                    "iterations": dn,
                                    "note": "state (u, u_head, p, g, mu = 24 B/px) is L2-resident at these sizes"}

    # ---------------- end-to-end timing through the host API (e2e)
    if not args.no_e2e:
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        out["e2e"] = {"value": world * n_upd * args.steps / e2e_s, "unit": UNIT,
                      "h2d_bytes_per_step": int(NF * W * H), "d2h_bytes_per_step": int(2 * frame_bytes),
                      "ms_per_step": e2e_s * 1e3 / args.steps,
                      "api": "rmd::Depthmap protocol: 8-bit gray frames (page-locked host memory, RMD_OPT_PINNED_INPUT) -> "
                             "SeedMatrix::setReferenceImage / update per frame -> final depth + convergence maps to host"}
        # the round-1 form for continuity: float frames from PAGEABLE memory (staged through the pinned ring)
        n_f32 = max(1, min(args.steps, 5))
        seeds.setReferenceImage(frames[0], poses[0], dmin, dmax)
        for k in range(1, NF):
            seeds.update(frames[k], poses[k])
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_f32):
            seeds.setReferenceImage(frames[0], poses[0], dmin, dmax)
            for k in range(1, NF):
                seeds.update(frames[k], poses[k])
            final_maps()
            gatherer.gather()
            host_maps.copy_(gatherer.packed, non_blocking=True)
            stream.synchronize()
        barrier()
        f32_s = max_over_ranks(time.perf_counter() - t0)
        out["e2e_f32_pageable"] = {"value": world * n_upd * n_f32 / f32_s, "unit": UNIT, "steps": n_f32,
                                   "h2d_bytes_per_step": int(NF * frame_bytes),
                                   "api": "SeedMatrix::update(float*) from pageable memory (reusable on return)"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        time.sleep(0.05)   # the library's ingest-copy helpers park after 0.5 ms without work
        out["cpu_baseline"] = cpu_baseline(args, frames, poses, dmin, dmax, seq)
    return out


# -------------------------------------------------------------- reference

def run_reference(args, rank, world, local_rank):
    """The reference's own path on this box.  Rank 0 alone runs it."""
    if rank != 0:
        return None
    import ref_binding as rb
    seq, frames_u8, frames, poses, dmin, dmax = load_sequence(args, 0)
    W, H, NF = args.width, args.height, args.frames
    n_upd = NF - 1
    frame_bytes = W * H * 4
    iters = args.denoise_iters
    use_cuda = False
    if not args.reference_cpu and rb.available(args.patch):
        try:
            use_cuda = rb.lib(args.patch).ref_device_count() > 0
        except OSError:
            use_cuda = False
    base_cfg = {"workload": workload_text(args, n_upd, per_gpu=False), "name": args.config}
    if use_cuda:
        ref = rb.RefSeeds(W, H, *seq.camera, patch=args.patch)
        rden = rb.RefDenoiser(W, H, patch=args.patch) if iters else None
        img32 = np.empty((H, W), np.float32)
        scale = np.float32(1.0 / 255.0)

        def step():
            # rmd::Depthmap::inputImage (src/depthmap.cpp:95-106): 8U -> 32F * (1/255.f) on the host, then SeedMatrix
            np.multiply(frames_u8[0], scale, out=img32)
            ref.set_reference(img32, poses[0], dmin, dmax)
            for k in range(1, NF):
                np.multiply(frames_u8[k], scale, out=img32)
                ref.update(img32, poses[k])
            if rden is not None:
                depth = rden.run(ref, dmax - dmin, 0.5, iters)      # downloadDenoisedDepthmap, src/depthmap.cpp:113
            else:
                ref.sync()
                depth = ref.download(0)
            return depth, ref.download(4)

        for _ in range(args.warmup):
            step()
        sampler = ClockSampler(0)
        sampler.start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
        v = n_upd * args.steps / dt
        # continuity with round 1: float frames handed to SeedMatrix directly (no host conversion)
        t0 = time.perf_counter()
        ref.set_reference(frames[0], poses[0], dmin, dmax)
        for k in range(1, NF):
            ref.update(frames[k], poses[k])
        ref.sync()
        ref.download(0), ref.download(4)
        v_f32 = n_upd / (time.perf_counter() - t0)
        return {"impl": "reference", "metric": args.metric, "value": v, "unit": UNIT, "n_gpus": 1,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": base_cfg,
                "cpu_baseline": {"value": v, "unit": UNIT, "kind": "reference", "cores": 1,
                                 "sample": "whole workload; NOTE the reference has no CPU implementation of this "
                                           "path -- this is its unmodified CUDA path (src/seed_matrix.cu and its "
                                           "kernels) rebuilt for sm_100a with texture objects "
                                           "(oracle/Makefile), 1 host thread driving 1 GPU, 8-bit host frames "
                                           "converted to float on the host (src/depthmap.cpp:105), sync pageable "
                                           "H2D per frame as the reference does"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": int(NF * frame_bytes),
                        "d2h_bytes_per_step": int(2 * frame_bytes)},
                "e2e_f32_pageable": {"value": v_f32, "unit": UNIT, "steps": 1,
                                     "api": "SeedMatrix::update(float*) without the host conversion"},
                "clocks": clocks}
    # CPU oracle port, bounded sample per step
    import oracle_binding as ob
    n = max(1, min(args.cpu_frames, n_upd))
    o = ob.OracleSeeds(W, H, *seq.camera, patch=args.patch)

    def step_cpu():
        o.set_reference(frames[0], poses[0], dmin, dmax)
        for k in range(1, n + 1):
            o.update(frames[k], poses[k])

    for _ in range(min(args.warmup, 1)):
        step_cpu()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        step_cpu()
    dt = time.perf_counter() - t0
    v = n * steps / dt
    return {"impl": "reference", "metric": args.metric, "value": v, "unit": UNIT, "n_gpus": 1, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3 / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": base_cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "kind": "port", "cores": ob.get_threads(),
                             "sample": f"frames 1..{n} of the sequence per step (search-heavy start), OpenMP"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def bind_to_gpu_numa_node(local_rank):
    """Keep this rank's host threads (and the pages they touch first: the pinned frames) on the NUMA node its GPU
    hangs off (8-GPU B200 hosts: GPU0-3 <-> node 0, GPU4-7 <-> node 1; SCALE_r01.json topology).  Best effort."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: whatever libraries print on fd 1
    # (e.g. "NCCL version ...") goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        out = run_reference(args, rank, world, local_rank)
        if out is not None:
            emit(out)
        return
    # ingest-copy helper threads of the library (pageable callers only): share the host's cores between the ranks
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    os.environ.setdefault("RMD_COPY_THREADS", str(max(3, min(15, (os.cpu_count() or 8) // (2 * max(1, local_world)) - 1))))
    numa = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        numa = bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
    out = run_ours(args, rank, world, local_rank)
    if rank == 0:
        out["config"]["host_numa_node_of_rank0"] = numa
        emit(out)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
