"""Bisect the K=8-vs-K=1 difference of the staged kernel: one update from the keyframe prior, three ways
(direct kernel, staged single, staged batched via updateMany of two identical keyframes); magnitudes of the
differences in matches and in the updated seeds (GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpg_open_remode_b200 as rmd
from rpg_open_remode_b200 import synth

W, H = 101, 77
seq = synth.SyntheticSequence(W, H, seed=0x5EED0010 + W)
cam = rmd.PinholeCamera(*seq.camera)
f0, f1, f2 = seq.frame(0), seq.frame(1, want_depth=False), seq.frame(2, want_depth=False)
dmin, dmax = float(f0.depth.min()), float(f0.depth.max())

def mk(variant):
    g = rmd.SeedMatrix(W, H, cam)
    g.setOption(rmd.OPT_KERNEL_VARIANT, variant)
    g.setOption(rmd.OPT_RECORD_MATCHES, 1)
    g.setReferenceImage(f0.image, f0.T_cam_world, dmin, dmax)
    return g

d, s, b1, b2 = mk(rmd.VARIANT_DIRECT), mk(rmd.VARIANT_STAGED), mk(rmd.VARIANT_STAGED), mk(rmd.VARIANT_STAGED)
for f in (f1, f2):
    d.update(f.image, f.T_cam_world)
    s.update(f.image, f.T_cam_world)
    rmd.SeedMatrix.updateMany([b1, b2], f.image, f.T_cam_world)
    snaps = {}
    for name, g in (("direct", d), ("single", s), ("batch0", b1), ("batch1", b2)):
        snaps[name] = dict(conv=g.downloadConvergence(), mu=g.downloadDepthmap(), s2=g.downloadSigmaSq(), a=g.downloadA(),
                           b=g.downloadB(), m=g.downloadEpipolarMatches())
    for other in ("single", "batch0", "batch1"):
        A, B = snaps["direct"], snaps[other]
        upd = (A["conv"] == 0) & (B["conv"] == 0)
        dm = np.abs(A["m"] - B["m"]).max(axis=2)
        line = f"frame {f.index} direct vs {other}: conv differ {int((A['conv'] != B['conv']).sum())}; "
        line += f"matches differ {int((dm[upd] != 0).sum())}/{int(upd.sum())} max {float(dm[upd].max()) if upd.any() else 0:.3e} px "
        line += f"(>0.35 px: {int((dm[upd] > 0.35).sum())}); "
        for k in ("mu", "s2", "a", "b"):
            x, y = A[k].astype(np.float64), B[k].astype(np.float64)
            rel = np.abs(x - y) / np.maximum(np.abs(x), 1e-30)
            line += f"{k} differ {int((A[k] != B[k]).sum())} max rel {float(rel.max()):.2e}; "
        print(line, flush=True)
        bad = np.argwhere((A["mu"] != B["mu"]))
        for (y, x) in bad[:4]:
            print(f"     ({x},{y}) match direct {A['m'][y, x]} {other} {B['m'][y, x]}  mu {A['mu'][y, x]:.7f} {B['mu'][y, x]:.7f} "
                  f"a {A['a'][y, x]:.5f} {B['a'][y, x]:.5f} conv {A['conv'][y, x]} {B['conv'][y, x]}")
