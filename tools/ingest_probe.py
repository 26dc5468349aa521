"""Device time of the ingest kernels (run under `ncu --metrics gpu__time_duration.sum`) and the end-to-end cost of
8-bit ingest with / without lens undistortion on the VGA sequence (GPU box)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpg_open_remode_b200 as rmd
from rpg_open_remode_b200 import synth

W, H, N = 640, 480, int(sys.argv[1]) if len(sys.argv) > 1 else 200
seq = synth.SyntheticSequence(W, H, seed=0x5EED0002)
fr = [seq.frame(k, want_depth=(k == 0)) for k in range(N)]
u8 = [f.image_u8 for f in fr]
dmin, dmax = float(fr[0].depth.min()), float(fr[0].depth.max())
for label, dist in (("u8", None), ("u8 + remap, identity lens (same filter work)", (0.0, 0.0, 0.0, 0.0)), ("u8 + remap, distorting lens (warps the synthetic frames: more filter work)", (-0.21, 0.06, 0.001, -0.0005))):
    g = rmd.SeedMatrix(W, H, rmd.PinholeCamera(*seq.camera))
    if dist:
        g.initUndistortionMap(*dist)
    best = 1e9
    for rep in range(3):
        g.setReferenceImage(u8[0], fr[0].T_cam_world, dmin, dmax)
        g.sync()
        t0 = time.perf_counter()
        for k in range(1, N):
            g.update(u8[k], fr[k].T_cam_world)
        g.sync()
        best = min(best, time.perf_counter() - t0)
    print(f"{label:75s}: {best * 1e3:.2f} ms for {N - 1} updates ({(N - 1) / best:.0f} fps)")
