"""Where and when do the product and the reference's CUDA kernels part ways over a long sequence?  (GPU box)
Runs the staged kernel, the direct kernel and oracle/_ref side by side and prints, every `every` frames:
staged == direct (must be exact), state agreement and the confusion matrix against the reference, the
image regions the disagreeing pixels lie in, and a sample of pixels whose state first differs.
    python tools/parity_diag.py [width height seed frames every patch]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import rpg_open_remode_b200 as rmd
from rpg_open_remode_b200 import synth
import ref_binding as rb

W, H, seed, N, every, patch = 1280, 720, 0x5EED0003, 260, 20, 5
argv = [int(a, 0) for a in sys.argv[1:]]
if len(argv) >= 2: W, H = argv[0], argv[1]
if len(argv) >= 3: seed = argv[2]
if len(argv) >= 4: N = argv[3]
if len(argv) >= 5: every = argv[4]
if len(argv) >= 6: patch = argv[5]
NAMES = ["UPDATE", "CONVERGED", "BORDER", "DIVERGED", "NO_MATCH", "NOT_VISIBLE"]
seq = synth.SyntheticSequence(W, H, seed=seed)
f0 = seq.frame(0)
dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
cam = rmd.PinholeCamera(*seq.camera)
gs = rmd.SeedMatrix(W, H, cam, patch_side=patch)
gd = rmd.SeedMatrix(W, H, cam, patch_side=patch)
gd.setOption(rmd.OPT_KERNEL_VARIANT, rmd.VARIANT_DIRECT)
r = rb.RefSeeds(W, H, *seq.camera, patch=patch)
for x in (gs, gd):
    x.setReferenceImage(f0.image, f0.T_cam_world, dmin, dmax)
r.set_reference(f0.image, f0.T_cam_world, dmin, dmax)
print(f"{W}x{H} seed {seed:#x} patch {patch}: depth {dmin:.3f}..{dmax:.3f}", flush=True)
prev_diff = np.zeros((H, W), bool)
for k in range(1, N):
    f = seq.frame(k, want_depth=False)
    for x in (gs, gd):
        x.update(f.image, f.T_cam_world)
    r.update(f.image, f.T_cam_world)
    if k % every and k != N - 1:
        continue
    cs, cd, cr = gs.downloadConvergence(), gd.downloadConvergence(), r.download(4)
    mus, mud, mur = gs.downloadDepthmap(), gd.downloadDepthmap(), r.download(0)
    sd_equal = bool(np.array_equal(cs, cd) and np.array_equal(mus, mud))
    diff = cs != cr
    conf = {}
    for a in range(5):
        for b in range(5):
            n = int(((cs == a) & (cr == b)).sum())
            if a != b and n:
                conf[f"ours {NAMES[a]} / ref {NAMES[b]}"] = n
    ys, xs = np.nonzero(diff)
    bands = np.histogram(xs, bins=4, range=(0, W))[0].tolist() if len(xs) else []
    rows = np.histogram(ys, bins=4, range=(0, H))[0].tolist() if len(ys) else []
    same = ~diff & (cr != 2)
    d = np.abs(mus.astype(np.float64) - mur)[same] / (dmax - dmin)
    print(f"frame {k:4d}: staged==direct {sd_equal}; states equal {1 - diff.mean():.5f} ({int(diff.sum())} differ); "
          f"mu bit-identical {float((d == 0).mean()):.4f}; converged ours {int((cs == 1).sum())} ref {int((cr == 1).sum())}; "
          f"x-quarters {bands} y-quarters {rows}; {conf}", flush=True)
    new = diff & ~prev_diff
    ny, nx = np.nonzero(new)
    if len(ny):
        a_o, b_o, s_o = gs.downloadA(), gs.downloadB(), gs.downloadSigmaSq()
        a_r, b_r, s_r = r.download(2), r.download(3), r.download(1)
        sel = np.random.default_rng(k).choice(len(ny), size=min(6, len(ny)), replace=False)
        for i in sel:
            y, x = int(ny[i]), int(nx[i])
            print(f"      ({x:4d},{y:3d}) ours {NAMES[cs[y, x]]:9s} mu {mus[y, x]:.5f} s2 {s_o[y, x]:.3e} a {a_o[y, x]:.3f} b {b_o[y, x]:.3f} | "
                  f"ref {NAMES[cr[y, x]]:9s} mu {mur[y, x]:.5f} s2 {s_r[y, x]:.3e} a {a_r[y, x]:.3f} b {b_r[y, x]:.3f} | gt {f0.depth[y, x]:.5f}")
    prev_diff = diff
