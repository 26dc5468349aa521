"""TV-L1 denoiser alone: device time per call and per iteration at VGA / 720p / 1080p, 50 and 200 iterations,
product vs the reference's own kernel (oracle/_ref) on the same seed state (GPU box).
`--once`: a single 720p / 50-iteration call (the target of the ncu capture, stage ncu_denoise)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import rpg_open_remode_b200 as rmd
from rpg_open_remode_b200 import synth
import ref_binding as rb

once = "--once" in sys.argv
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
for (W, H) in ((1280, 720),) if once else ((640, 480), (1280, 720), (1920, 1080)):
    seq = synth.SyntheticSequence(W, H, seed=0x5EED0003)
    f0 = seq.frame(0)
    dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
    g = rmd.SeedMatrix(W, H, rmd.PinholeCamera(*seq.camera))
    g.setStream(stream.cuda_stream)
    g.setReferenceImage(f0.image, f0.T_cam_world, dmin, dmax)
    for k in range(1, 25):
        f = seq.frame(k, want_depth=False)
        g.update(f.image, f.T_cam_world)
    den = rmd.DepthmapDenoiser(W, H)
    den.setStream(stream.cuda_stream)
    den.setLargeSigmaSq(dmax - dmin)
    out = torch.empty((H, W), dtype=torch.float32, device=dev)
    if once:
        den.denoiseSeedsToDevice(g, out.data_ptr(), W * 4, 0.5, 50)
        torch.cuda.synchronize()
        den.denoiseSeedsToDevice(g, out.data_ptr(), W * 4, 0.5, 50)
        torch.cuda.synchronize()
        break
    r = rden = None
    if rb.available(5):
        r = rb.RefSeeds(W, H, *seq.camera)
        r.set_reference(f0.image, f0.T_cam_world, dmin, dmax)
        for fid, a in ((0, g.downloadDepthmap()), (1, g.downloadSigmaSq()), (2, g.downloadA()), (3, g.downloadB())):
            r.upload(fid, a)
        rden = rb.RefDenoiser(W, H)
    for iters in (50, 200):
        for _ in range(3):
            den.denoiseSeedsToDevice(g, out.data_ptr(), W * 4, 0.5, iters)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a.record(stream)
        for _ in range(reps):
            den.denoiseSeedsToDevice(g, out.data_ptr(), W * 4, 0.5, iters)
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            host = den.denoiseSeeds(g, 0.5, iters)
        host_ms = (time.perf_counter() - t0) / reps * 1e3
        line = (f"{W}x{H} {iters:3d} it: product device {ms:7.3f} ms ({ms * 1e3 / iters:5.2f} us/it, "
                f"{40.0 * W * H * iters / (ms * 1e-3) / 1e9:7.0f} GB/s algorithmic), host call {host_ms:7.3f} ms")
        if rden is not None:
            rden.run(r, dmax - dmin, 0.5, iters)
            t0 = time.perf_counter()
            for _ in range(reps):
                ref_out = rden.run(r, dmax - dmin, 0.5, iters)
            ref_ms = (time.perf_counter() - t0) / reps * 1e3
            d = np.abs(host - ref_out) / (dmax - dmin)
            line += f"; reference host call {ref_ms:7.3f} ms; |diff|/range median {np.median(d):.1e} p99 {np.percentile(d, 99):.1e}"
        print(line, flush=True)
