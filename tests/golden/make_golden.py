"""Generates tests/golden/ref_cuda_96x72.npz from the REFERENCE's own CUDA kernels
(oracle/_ref/librmd_ref.so = /root/reference/src/*.cu rebuilt for sm_100a, recipe
oracle/Makefile).  Must run on a GPU box:  python tests/golden/make_golden.py

Contents: the inputs (8-bit frames, poses, depth range, camera) of a 96x72, 9-frame
synthetic sequence and the reference's outputs after seedInit, after update 1 and
after update 8 (mu, sigma_sq, a, b, convergence, sum_templ, const_templ_denom,
epipolar matches of update 1), plus its 50-iteration denoised map.
CPU tests pin the oracle to these vectors; GPU tests pin the product to them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_binding as rb  # noqa: E402
from rpg_open_remode_b200 import synth  # noqa: E402

W, H, N = 96, 72, 9
seq = synth.SyntheticSequence(W, H, seed=0x5EED0601)
frames = [seq.frame(k, want_depth=(k == 0)) for k in range(N)]
f0 = frames[0]
dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
r = rb.RefSeeds(W, H, *seq.camera, patch=5)
r.set_reference(f0.image, f0.T_cam_world, dmin, dmax)
out = {
    "width": W, "height": H, "camera": np.array(seq.camera, np.float32), "min_depth": np.float32(dmin),
    "max_depth": np.float32(dmax),
    "frames_u8": np.stack([f.image_u8 for f in frames]),
    "T_cam_world": np.stack([f.T_cam_world for f in frames]).astype(np.float32),
    "depth0": f0.depth,
    "init_mu": r.download(0), "init_sigma_sq": r.download(1), "init_a": r.download(2), "init_b": r.download(3),
    "sum_templ": r.download(5), "const_templ_denom": r.download(6),
}
for k in range(1, N):
    r.update(frames[k].image, frames[k].T_cam_world)
    if k in (1, 8):
        out[f"u{k}_mu"] = r.download(0)
        out[f"u{k}_sigma_sq"] = r.download(1)
        out[f"u{k}_a"] = r.download(2)
        out[f"u{k}_b"] = r.download(3)
        out[f"u{k}_conv"] = r.download(4).astype(np.int8)
        if k == 1:
            out["u1_matches"] = r.download(7)
        out[f"u{k}_converged_count"] = np.int64(r.converged_count())
        out[f"u{k}_dist_from_ref"] = np.float32(r.dist_from_ref())
den = rb.RefDenoiser(W, H)
out["denoised_50"] = den.run(r, dmax - dmin, 0.5, 50)
np.savez_compressed(os.path.join(HERE, "ref_cuda_96x72.npz"), **out)
print("wrote", os.path.join(HERE, "ref_cuda_96x72.npz"), {k: getattr(v, "shape", None) for k, v in out.items()})
