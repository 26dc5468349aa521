"""Secondary measurements for the other BASELINE configs (GPU box), next to the
reference's own CUDA build where it can run them:

  * depth filter, device-resident frames: 1280x720 (5x5) and 1920x1080 (7x7), first 60 frames
  * TV-L1 denoiser: VGA and 720p, 50 and 200 iterations (device time, CUDA events)
  * reductions: countEqual / sum on 752x480
  * frame ingest: float vs 8-bit host frames through update()

Prints one JSON object; tools/gpu_trip.sh (stage `extra`) stores it under gpurun_out/ and the
summary lives in profiles/r01_extra_bench.md.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import ref_binding as rb  # noqa: E402
import rpg_open_remode_b200 as rmd  # noqa: E402
from rpg_open_remode_b200 import synth  # noqa: E402

out = {}
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)


def render(W, H, n, seed):
    seq = synth.SyntheticSequence(W, H, seed=seed)
    frames = np.empty((n, H, W), np.float32)
    poses = np.empty((n, 12), np.float32)
    for k in range(n):
        f = seq.frame(k, want_depth=(k == 0))
        frames[k] = f.image
        poses[k] = f.T_cam_world.reshape(12)
        if k == 0:
            dmin, dmax = float(f.depth.min()), float(f.depth.max())
    return seq, frames, poses, dmin, dmax


def filter_fps(W, H, patch, n, seed, label):
    seq, frames, poses, dmin, dmax = render(W, H, n, seed)
    g = rmd.SeedMatrix(W, H, rmd.PinholeCamera(*seq.camera), patch_side=patch)
    g.setStream(stream.cuda_stream)
    d_frames = torch.from_numpy(frames).to(dev)
    best = None
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.setReferenceImageDevice(d_frames[0].data_ptr(), W * 4, poses[0], dmin, dmax)
        e0.record(stream)
        g.updateDeviceBatch(d_frames[1].data_ptr(), W * H * 4, W * 4, poses[1:])
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    res = {"frames": n - 1, "ms": best, "fps": (n - 1) / best * 1e3,
           "achieved_GBps_52B_per_px": 52.0 * W * H * (n - 1) / (best * 1e-3) / 1e9}
    if rb.available(patch):
        r = rb.RefSeeds(W, H, *seq.camera, patch=patch)
        tb = None
        for rep in range(2):
            r.set_reference(frames[0], poses[0], dmin, dmax)
            r.sync()
            t0 = time.perf_counter()
            for k in range(1, n):
                r.update(frames[k], poses[k])
            r.sync()
            dt = time.perf_counter() - t0
            tb = dt if tb is None else min(tb, dt)
        res["reference_cuda_fps_host_frames"] = (n - 1) / tb
    out[label] = res
    return g, seq, dmin, dmax


g720, seq720, dmin720, dmax720 = filter_fps(1280, 720, 5, 60, 0x5EED0003, "filter_720p_5x5_first60")
filter_fps(1920, 1080, 7, 40, 0x5EED0004, "filter_1080p_7x7_first40")
gvga, seqvga, dminv, dmaxv = filter_fps(640, 480, 5, 60, 0x5EED0002, "filter_vga_5x5_first60")


def denoise_ms(g, W, H, rng, iters, label):
    den = rmd.DepthmapDenoiser(W, H)
    den.setStream(stream.cuda_stream)
    den.setLargeSigmaSq(rng)
    dst = torch.empty((H, W), dtype=torch.float32, device=dev)
    best = None
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        den.denoiseSeedsToDevice(g, dst.data_ptr(), W * 4, 0.5, iters)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    t0 = time.perf_counter()
    host = den.denoiseSeeds(g, 0.5, iters)
    wall = (time.perf_counter() - t0) * 1e3
    out[label] = {"iterations": iters, "device_ms": best, "us_per_iteration": best / max(iters, 1) * 1e3,
                  "achieved_GBps_40B_per_px_iter": 40.0 * W * H * iters / (best * 1e-3) / 1e9,
                  "host_call_ms_incl_d2h": wall, "finite": bool(np.isfinite(host).all())}


for it in (50, 200):
    denoise_ms(gvga, 640, 480, dmaxv - dminv, it, f"denoise_vga_{it}")
    denoise_ms(g720, 1280, 720, dmax720 - dmin720, it, f"denoise_720p_{it}")

if rb.available(5):
    seq, frames, poses, dmin, dmax = render(640, 480, 16, 0x5EED0002)
    r = rb.RefSeeds(640, 480, *seq.camera)
    r.set_reference(frames[0], poses[0], dmin, dmax)
    for k in range(1, 16):
        r.update(frames[k], poses[k])
    rd = rb.RefDenoiser(640, 480)
    for it in (50, 200):
        rd.run(r, dmax - dmin, 0.5, it)
        t0 = time.perf_counter()
        rd.run(r, dmax - dmin, 0.5, it)
        out[f"reference_cuda_denoise_vga_{it}_host_call_ms"] = (time.perf_counter() - t0) * 1e3

# reductions
rng = np.random.default_rng(1)
ints = rng.integers(0, 256, size=(480, 752), dtype=np.int32)
di = rmd.DeviceImage(752, 480, "int32")
di.setDevData(ints)
red = rmd.ImageReducer("int32")
red.countEqual(di, 2)
t0 = time.perf_counter()
for _ in range(200):
    red.countEqual(di, 2)
out["count_equal_752x480_host_call_us"] = (time.perf_counter() - t0) / 200 * 1e6
if rb.available(5):
    rb.reduce_count_eq(ints, 2)
    t0 = time.perf_counter()
    for _ in range(20):
        rb.reduce_count_eq(ints, 2)
    out["reference_cuda_count_equal_host_call_us_incl_upload"] = (time.perf_counter() - t0) / 20 * 1e6

# ingest: float vs u8 host frames
seq, frames, poses, dmin, dmax = render(640, 480, 200, 0x5EED0002)
u8 = np.rint(frames * 255.0).astype(np.uint8)
for label, src in (("e2e_float_frames_fps", frames), ("e2e_u8_frames_fps", u8)):
    g = rmd.SeedMatrix(640, 480, rmd.PinholeCamera(*seq.camera))
    best = None
    for rep in range(3):
        g.setReferenceImage(src[0], poses[0], dmin, dmax)
        g.sync()
        t0 = time.perf_counter()
        for k in range(1, 200):
            g.update(src[k], poses[k])
        g.sync()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out[label] = 199 / best

# ingest with lens undistortion: the px4 launch-file camera, 752x480
# (a) the whole 8-bit ingest chain of rmd::Depthmap::inputImage on the CPU with OpenCV, as the reference runs it;
# (b) the same frame through the library: H2D of the 8-bit frame + one fused remap/convert kernel (+ D2H for the probe);
# (c) VGA sequence end to end with distorted 8-bit frames in.
try:
    import cv2
    Wd, Hd = 752, 480
    camd = (418.715779404372, 418.186411716775, 397.573670639476, 246.235858293295)
    distd = (-0.294854287021541, 0.0780596214365872, -0.000520874224877783, 9.42576963868232e-06)
    K = np.array([[camd[0], 0, camd[2]], [0, camd[1], camd[3]], [0, 0, 1]], np.float32)
    m1, m2 = cv2.initUndistortRectifyMap(K, np.array([distd], np.float32), np.eye(3), K, (Wd, Hd), cv2.CV_16SC2)
    imgs = [rng.integers(0, 256, size=(Hd, Wd), dtype=np.uint8) for _ in range(50)]
    cv2.setNumThreads(0)
    t0 = time.perf_counter()
    for k in range(200):
        und = cv2.remap(imgs[k % 50], m1, m2, cv2.INTER_LINEAR)
        flt = cv2.multiply(und, 1.0, scale=1.0 / 255.0, dtype=cv2.CV_32F)
    out["ingest_752x480_opencv_cpu_1thread_us_per_frame"] = (time.perf_counter() - t0) / 200 * 1e6
    gd = rmd.SeedMatrix(Wd, Hd, rmd.PinholeCamera(camd[0], camd[1], camd[2], camd[3]))
    gd.initUndistortionMap(*distd)
    got = gd.undistort(imgs[0])
    out["ingest_752x480_matches_opencv"] = bool(np.array_equal(got, cv2.remap(imgs[0], m1, m2, cv2.INTER_LINEAR)))
    t0 = time.perf_counter()
    for k in range(200):
        gd.undistort(imgs[k % 50])
    out["ingest_752x480_library_h2d_kernel_d2h_sync_us_per_frame"] = (time.perf_counter() - t0) / 200 * 1e6
    # device time of the fused remap+convert kernel: set_reference_u8 = H2D + that kernel + seed init; use CUDA events
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gd.setStream(stream.cuda_stream)
    gd.setReferenceImage(imgs[0], poses[0], dmin, dmax)
    for k in range(1, 4):
        gd.update(imgs[k], poses[k])
    gd.sync()
except Exception as e:   # cv2 missing on the box: report it, do not fail the whole run
    out["ingest_752x480_error"] = repr(e)

# identity lens: the remap kernel runs, the frames (hence the filter work) stay the same
g.initUndistortionMap(0.0, 0.0, 0.0, 0.0)
best = None
for rep in range(3):
    g.setReferenceImage(u8[0], poses[0], dmin, dmax)
    g.sync()
    t0 = time.perf_counter()
    for k in range(1, 200):
        g.update(u8[k], poses[k])
    g.sync()
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
out["e2e_u8_frames_with_undistortion_fps"] = 199 / best
g.clearUndistortionMap()

# point cloud: VGA keyframe after 199 updates
import oracle_binding as ob  # noqa: E402  (the reference's CPU loop, transcribed: the baseline of this row)
g.setReferenceImage(u8[0], poses[0], dmin, dmax)
for k in range(1, 200):
    g.update(u8[k], poses[k])
g.sync()
pts, n_pts = g.pointCloud()
t0 = time.perf_counter()
for _ in range(20):
    pts, n_pts = g.pointCloud()
out["point_cloud_vga_points"] = int(n_pts)
out["point_cloud_vga_library_ms_incl_d2h_of_points"] = (time.perf_counter() - t0) / 20 * 1e3
t0 = time.perf_counter()
for _ in range(20):
    mu, conv = g.downloadDepthmap(), g.downloadConvergence()
t_dl = (time.perf_counter() - t0) / 20
T_world_ref = ob.se3_inv(poses[0].reshape(3, 4))
camf = [float(np.float32(c)) for c in seq.camera]
t0 = time.perf_counter()
for _ in range(5):
    ref_pts = ob.point_cloud(mu, conv, u8[0], *camf, T_world_ref)
t_cpu = (time.perf_counter() - t0) / 5
out["point_cloud_vga_reference_way_ms"] = {"download_depth_and_convergence": t_dl * 1e3,
                                            "cpu_loop_two_passes_count_then_fill": t_cpu * 1e3}
out["point_cloud_vga_matches_cpu_loop"] = bool(np.array_equal(pts, ref_pts))

print(json.dumps(out, indent=1))
