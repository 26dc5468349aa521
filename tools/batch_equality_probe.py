"""Debug probe of the batched (multi-keyframe) staged launch: K keyframes through SeedMatrix.updateMany vs the same
keyframes one by one, reporting the first frame and the tiles where they part (GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpg_open_remode_b200 as rmd
from rpg_open_remode_b200 import synth

W, H, N = 320, 240, 14
mode = sys.argv[1] if len(sys.argv) > 1 else "same"      # same: K copies of one keyframe; diff: different start frames
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
seq = synth.SyntheticSequence(W, H, seed=0x5EED0001)
cam = rmd.PinholeCamera(*seq.camera)
fr = [seq.frame(k, want_depth=(k == 0)) for k in range(N)]
dmin, dmax = float(fr[0].depth.min()), float(fr[0].depth.max())
starts = [0] * K if mode == "same" else list(range(0, 2 * K, 2))
batch = [rmd.SeedMatrix(W, H, cam) for _ in range(K)]
alone = [rmd.SeedMatrix(W, H, cam) for _ in range(K)]
for k in range(N):
    live = [i for i in range(K) if k > starts[i]]
    if live:
        rmd.SeedMatrix.updateMany([batch[i] for i in live], fr[k].image, fr[k].T_cam_world)
        for i in live:
            alone[i].update(fr[k].image, fr[k].T_cam_world)
    for i in range(K):
        if k == starts[i]:
            batch[i].setReferenceImage(fr[k].image, fr[k].T_cam_world, dmin, dmax)
            alone[i].setReferenceImage(fr[k].image, fr[k].T_cam_world, dmin, dmax)
    for i in live:
        cb, ca = batch[i].downloadConvergence(), alone[i].downloadConvergence()
        mb, ma = batch[i].downloadDepthmap(), alone[i].downloadDepthmap()
        bad = (cb != ca) | (mb != ma)
        if bad.any():
            ys, xs = np.nonzero(bad)
            tiles = sorted(set((int(y) // 8) * ((W + 31) // 32) + int(x) // 32 for y, x in zip(ys, xs)))
            print(f"frame {k} keyframe {i} (live {live}): {int(bad.sum())} pixels differ in {len(tiles)} tiles, first tiles {tiles[:12]}; "
                  f"states batch {np.bincount(cb.ravel(), minlength=5).tolist()} alone {np.bincount(ca.ravel(), minlength=5).tolist()}; "
                  f"counts batch {batch[i].getConvergedCount()} alone {alone[i].getConvergedCount()}", flush=True)
        else:
            print(f"frame {k} keyframe {i} (live {live}): equal", flush=True)
