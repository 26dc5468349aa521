"""Seed-major mode (csrc/depth_filter_seeds.cu): once few seeds are still updated the handle keeps them as a compact
list and a launch walks every listed seed through its frames warp by warp.  It is an organisation of the same
per-seed steps, so everything here is compared BIT FOR BIT with the tile-organised kernel (seed mode off) and the
direct kernel: entered at different times (threshold 100 % = as soon as the first statistics arrive, 30 %, default),
through the per-frame host path and through multi-frame device batches, with keyframe sets updated together, after a
state upload in the middle (the list is dropped and rebuilt), for 5x5 and 7x7 patches; getConvergedCount keeps
counting."""
import numpy as np
import pytest

import rpg_open_remode_b200 as rmd
from rpg_open_remode_b200 import synth

pytestmark = pytest.mark.gpu
FIELDS = ("conv", "mu", "sigma_sq", "a", "b")


def _snap(g):
    return {"conv": g.downloadConvergence(), "mu": g.downloadDepthmap(), "sigma_sq": g.downloadSigmaSq(),
            "a": g.downloadA(), "b": g.downloadB()}


def _same(A, B, what):
    for name in FIELDS:
        assert np.array_equal(A[name], B[name]), f"{what}: {name} differs at {(A[name] != B[name]).sum()} pixels"


def _new(seq, f0, dmin, dmax, variant=rmd.VARIANT_STAGED, pct=None, patch=5):
    g = rmd.SeedMatrix(seq.width, seq.height, rmd.PinholeCamera(*seq.camera), patch_side=patch)
    g.setOption(rmd.OPT_KERNEL_VARIANT, variant)
    if pct is not None:
        g.setOption(rmd.OPT_SEED_MODE_PCT, pct)
    g.setReferenceImage(f0.image, f0.T_cam_world, dmin, dmax)
    return g


@pytest.mark.parametrize("patch,size,n", [(5, (320, 240), 90), (7, (203, 131), 40)])
def test_seed_mode_equals_tile_mode_and_direct(patch, size, n):
    W, H = size
    seq = synth.SyntheticSequence(W, H, seed=0x5EED0030 + W)
    frames = [seq.frame(k, want_depth=(k == 0)) for k in range(n)]
    f0 = frames[0]
    dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
    ref = _new(seq, f0, dmin, dmax, rmd.VARIANT_DIRECT, patch=patch)
    tile = _new(seq, f0, dmin, dmax, pct=0, patch=patch)
    early = _new(seq, f0, dmin, dmax, pct=100, patch=patch)     # seed-major from the third or fourth frame on
    mid = _new(seq, f0, dmin, dmax, pct=30, patch=patch)
    dflt = _new(seq, f0, dmin, dmax, pct=8, patch=patch)
    for k in range(1, n):
        for g in (ref, tile, early, mid, dflt):
            g.update(frames[k].image, frames[k].T_cam_world)
        if k in (3, 8, 20, n // 2, n - 1):
            R = _snap(ref)
            for name, g in (("tile", tile), ("seed-major from the start", early), ("seed-major at 30 %", mid), ("seed-major at 8 %", dflt)):
                _same(_snap(g), R, f"{name}, frame {k}")
                assert g.getConvergedCount() == ref.getConvergedCount() == int((R["conv"] == 1).sum()), (name, k)
    assert early.launchCount()[0] == n - 1


def test_seed_mode_device_batches_and_state_upload():
    import torch
    W, H, N = 320, 240, 70
    seq = synth.SyntheticSequence(W, H, seed=0x5EED0031)
    frames = [seq.frame(k, want_depth=(k == 0)) for k in range(N)]
    f0 = frames[0]
    dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
    poses = np.stack([f.T_cam_world.reshape(12) for f in frames]).astype(np.float32)
    dense = torch.from_numpy(np.stack([f.image for f in frames])).to(torch.device("cuda", 0))
    want = _new(seq, f0, dmin, dmax, pct=0)
    for k in range(1, N):
        want.update(frames[k].image, poses[k])
    Wn = _snap(want)
    # device batches: chained tile launches until the statistics arrive, then seed-major launches of up to 16 frames
    g = _new(seq, f0, dmin, dmax, pct=100)
    g.setOption(rmd.OPT_CHAIN_FRAMES, 8)
    g.updateDeviceBatch(dense[1].data_ptr(), W * H * 4, W * 4, poses[1:7])
    g.updateDeviceBatch(dense[7].data_ptr(), W * H * 4, W * 4, poses[7:50])
    g.updateDeviceBatch(dense[50].data_ptr(), W * H * 4, W * 4, poses[50:])
    _same(_snap(g), Wn, "device batches, seed-major")
    assert g.getConvergedCount() == want.getConvergedCount()
    assert g.launchCount()[0] < (N - 1) // 3      # far fewer launches than frames
    g.sync()
    # a state upload in the middle drops the list; the tile work list is rebuilt, seed-major mode is entered again
    h = _new(seq, f0, dmin, dmax, pct=100)
    ref = _new(seq, f0, dmin, dmax, rmd.VARIANT_DIRECT)
    for k in range(1, 30):
        h.update(frames[k].image, poses[k])
        ref.update(frames[k].image, poses[k])
    state = _snap(ref)
    for fid, name in ((rmd.FIELD_MU, "mu"), (rmd.FIELD_SIGMA_SQ, "sigma_sq"), (rmd.FIELD_A, "a"), (rmd.FIELD_B, "b"),
                      (rmd.FIELD_CONVERGENCE, "conv")):
        h.uploadState(fid, state[name])
        ref.uploadState(fid, state[name])
    for k in range(30, N):
        h.update(frames[k].image, poses[k])
        ref.update(frames[k].image, poses[k])
    _same(_snap(h), _snap(ref), "after a state upload")
    assert h.getConvergedCount() == ref.getConvergedCount()
    torch.cuda.synchronize()


def test_seed_mode_with_keyframe_sets():
    """Keyframes of one set enter seed-major mode at different times; rmd_seeds_update_many batches the ones still
    tile-organised and runs the others one by one -- same results as updating each keyframe alone."""
    W, H, N = 320, 240, 60
    seq = synth.SyntheticSequence(W, H, seed=0x5EED0001)
    frames = [seq.frame(k, want_depth=(k == 0)) for k in range(N)]
    dmin, dmax = float(frames[0].depth.min()), float(frames[0].depth.max())
    cam = rmd.PinholeCamera(*seq.camera)
    starts, pcts = [0, 3, 7], [100, 30, 0]
    batch = [rmd.SeedMatrix(W, H, cam) for _ in starts]
    alone = [rmd.SeedMatrix(W, H, cam) for _ in starts]
    for g, pct in zip(batch, pcts):
        g.setOption(rmd.OPT_SEED_MODE_PCT, pct)
    for g in alone:
        g.setOption(rmd.OPT_SEED_MODE_PCT, 0)
    for k in range(N):
        live = [i for i, s0 in enumerate(starts) if k > s0]
        if live:
            rmd.SeedMatrix.updateMany([batch[i] for i in live], frames[k].image, frames[k].T_cam_world)
            for i in live:
                alone[i].update(frames[k].image, frames[k].T_cam_world)
        for i, s0 in enumerate(starts):
            if k == s0:
                batch[i].setReferenceImage(frames[k].image, frames[k].T_cam_world, dmin, dmax)
                alone[i].setReferenceImage(frames[k].image, frames[k].T_cam_world, dmin, dmax)
    for i in range(len(starts)):
        _same(_snap(batch[i]), _snap(alone[i]), f"keyframe {i}")
        assert batch[i].getConvergedCount() == alone[i].getConvergedCount()
