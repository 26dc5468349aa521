"""RMD_OPT_HOST_FRAME_GROUP: host frames are uploaded when they are handed over, their kernel is deferred until a
group is full and then ONE chained launch covers the group.  Nothing observable may change: every query in the middle
of a group (converged count, downloads, distance from the reference, the denoiser, a new reference frame, a state
upload, a device-side update) first launches what is waiting.  Everything is compared BIT FOR BIT with group 1."""
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


def _sequence(W, H, n, seed):
    seq = synth.SyntheticSequence(W, H, seed=seed)
    frames = [seq.frame(k, want_depth=(k == 0)) for k in range(n)]
    return seq, frames, float(frames[0].depth.min()), float(frames[0].depth.max())


def _new(seq, group, patch=5):
    g = rmd.SeedMatrix(seq.width, seq.height, rmd.PinholeCamera(*seq.camera), patch_side=patch)
    g.setOption(rmd.OPT_HOST_FRAME_GROUP, group)
    return g


@pytest.mark.parametrize("patch,size,n", [(5, (320, 240), 45), (7, (203, 131), 30)])
@pytest.mark.parametrize("u8", [False, True])
def test_grouped_host_frames_equal_per_frame_launches(patch, size, n, u8):
    W, H = size
    seq, frames, dmin, dmax = _sequence(W, H, n, 0x5EED0040 + W)
    img = (lambda f: f.image_u8) if u8 else (lambda f: f.image)
    handles = {grp: _new(seq, grp, patch) for grp in (1, 2, 5, 8)}
    for g in handles.values():
        g.setReferenceImage(img(frames[0]), frames[0].T_cam_world, dmin, dmax)
    one = handles[1]
    for k in range(1, n):
        for g in handles.values():
            g.update(img(frames[k]), frames[k].T_cam_world)
        if k in (3, 11, n - 1):                  # in the middle of a group for some group size
            want = _snap(one)
            for grp, g in handles.items():
                assert g.getDistFromRef() == one.getDistFromRef()
                assert g.getConvergedCount() == one.getConvergedCount(), (grp, k)
                _same(_snap(g), want, f"group {grp}, frame {k}")
    fused = {grp: g.launchCount()[0] for grp, g in handles.items()}
    assert fused[1] == n - 1
    assert fused[8] < fused[2] < fused[1]        # the launch count is the only thing that changes


def test_grouped_frames_survive_interleaved_calls():
    import torch
    W, H, N = 320, 240, 40
    seq, frames, dmin, dmax = _sequence(W, H, N, 0x5EED0041)
    dev = torch.device("cuda", 0)
    dense = torch.from_numpy(np.stack([f.image for f in frames])).to(dev)
    poses = np.stack([f.T_cam_world.reshape(12) for f in frames]).astype(np.float32)
    den = {grp: rmd.DepthmapDenoiser(W, H) for grp in (1, 8)}
    out = {}
    for grp in (1, 8):
        g = _new(seq, grp)
        den[grp].setLargeSigmaSq(dmax - dmin)
        g.setReferenceImage(frames[0].image, frames[0].T_cam_world, dmin, dmax)
        for k in range(1, 6):
            g.update(frames[k].image, poses[k])
        # device-side frames in the middle of a group: the waiting host frames come first
        g.updateDevice(dense[6].data_ptr(), W * 4, poses[6])
        g.updateDeviceBatch(dense[7].data_ptr(), W * H * 4, W * 4, poses[7:12])
        for k in range(12, 15):
            g.update(frames[k].image, poses[k])
        # a state upload in the middle of a group
        mu = g.downloadDepthmap()
        g.uploadState(rmd.FIELD_MU, mu)
        for k in range(15, 21):
            g.update(frames[k].image, poses[k])
        denoised = den[grp].denoiseSeeds(g, 0.5, 16)           # reads the seeds: the group is launched first
        # a new keyframe while three frames of the old one wait
        for k in range(21, 24):
            g.update(frames[k].image, poses[k])
        before = _snap(g)
        g.setReferenceImage(frames[24].image, frames[24].T_cam_world, dmin, dmax)
        for k in range(25, N):
            g.update(frames[k].image, poses[k])
        out[grp] = (before, denoised, _snap(g), g.getConvergedCount())
        g.sync()
    _same(out[8][0], out[1][0], "before the new reference")
    assert np.array_equal(out[8][1], out[1][1]), "denoised map"
    _same(out[8][2], out[1][2], "end of the second keyframe")
    assert out[8][3] == out[1][3]


def test_group_option_range():
    seq, frames, dmin, dmax = _sequence(160, 120, 1, 0x5EED0042)
    g = rmd.SeedMatrix(160, 120, rmd.PinholeCamera(*seq.camera))
    for bad in (0, 9, -1):
        with pytest.raises(rmd.RmdError):
            g.setOption(rmd.OPT_HOST_FRAME_GROUP, bad)
    g.setOption(rmd.OPT_HOST_FRAME_GROUP, 8)
    g.setOption(rmd.OPT_HOST_FRAME_GROUP, 1)
