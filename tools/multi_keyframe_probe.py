"""Aggregate throughput of K live keyframes fed by one VGA frame stream through rmd_seeds_update_many
(SURVEY 8f row 2): one upload per frame, K fused kernels on K streams (GPU box)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpg_open_remode_b200 as rmd
from rpg_open_remode_b200 import synth, node

W, H, N = 640, 480, 200
seq = synth.SyntheticSequence(W, H, seed=0x5EED0002)
fr = [seq.frame(k, want_depth=(k == 0)) for k in range(N)]
dmin, dmax = float(fr[0].depth.min()), float(fr[0].depth.max())
cam = rmd.PinholeCamera(*seq.camera)
for K, warp_tiles in ((1, 8), (2, 8), (4, 8), (8, 8), (8, 0)):
    ks = node.KeyframeSet(W, H, cam, n=K)
    for sd in ks.seeds:
        sd.setOption(rmd.OPT_TUNE_WARP_TILE_SEEDS, warp_tiles)
    best = 1e9
    for rep in range(2):
        for i in range(K):
            ks.setReferenceImage(i, fr[0].image, fr[0].T_cam_world, dmin, dmax)
        for s in ks.seeds:
            s.sync()
        t0 = time.perf_counter()
        for k in range(1, N):
            ks.update(fr[k].image, fr[k].T_cam_world)
        for s in ks.seeds:
            s.sync()
        best = min(best, time.perf_counter() - t0)
    print(f"K = {K} (warp tiles {warp_tiles}): {best * 1e3:7.2f} ms for {N - 1} frames -> {(N - 1) / best:8.0f} frames/s, "
          f"{K * (N - 1) / best:8.0f} keyframe-updates/s ({best * 1e3 / K:.2f} ms per keyframe)", flush=True)
