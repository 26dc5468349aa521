"""Point-cloud extraction -- SURVEY.md 8f row 3: rmd::Publisher::publishPointCloud (src/publisher.cpp:54-86).

In the reference this step is CPU code (a double loop over the downloaded maps), so the oracle
(oracle/rmd_oracle_pointcloud.c) is a transcription of it.  Point ORDER, COUNT and INTENSITY are index /
integer work: exact.  The coordinates are fp32; the product evaluates them with IEEE round-to-nearest
intrinsics in the reference's operation order, so the bar here is also bit-exact against the oracle.

  * CPU: the oracle against an independent numpy-float32 evaluation of the same formula, and
    size-independent properties (distance to the camera centre = depth, row-major order, count).
  * GPU: the product (C-ABI rmd_seeds_point_cloud) against the oracle on real filter output, from the
    seeds' depth and from a denoised device image; the truncation contract of `capacity`.
"""
import numpy as np
import pytest

import oracle_binding as ob

F = np.float32


def _numpy_points(depth, conv, ref_u8, fx, fy, cx, cy, T):
    """The formula of src/publisher.cpp:73-83 in numpy float32, one rounding per operation."""
    h, w = depth.shape
    yy, xx = np.mgrid[0:h, 0:w]
    vx = (xx.astype(F) - F(cx)) / F(fx)
    vy = (yy.astype(F) - F(cy)) / F(fy)
    vz = np.ones_like(vx)
    inv = F(1) / np.sqrt((vx * vx + vy * vy) + vz * vz)
    px, py, pz = (vx * inv) * depth, (vy * inv) * depth, (vz * inv) * depth
    T = np.asarray(T, F).reshape(3, 4)
    out = [((T[r, 0] * px + T[r, 1] * py) + T[r, 2] * pz) + T[r, 3] for r in range(3)]
    sel = conv == 1
    return np.stack([out[0][sel], out[1][sel], out[2][sel], ref_u8[sel].astype(F)], axis=1)


def _random_case(w, h, seed):
    rng = np.random.default_rng(seed)
    depth = rng.uniform(0.5, 3.0, (h, w)).astype(F)
    conv = rng.integers(0, 6, (h, w)).astype(np.int32)
    ref = rng.integers(0, 256, (h, w)).astype(np.uint8)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    T = ob.se3_from_quat(*[float(v) for v in q], 0.3, -1.2, 2.5)
    return depth, conv, ref, T


@pytest.mark.parametrize("size", [(7, 5), (160, 120), (752, 480)])
def test_oracle_point_cloud_equals_numpy_float32(size):
    w, h = size
    depth, conv, ref, T = _random_case(w, h, 100 + w)
    cam = (481.2 * w / 640, -480.0 * h / 480, (w - 1) / 2, (h - 1) / 2)
    cam = tuple(float(F(c)) for c in cam)
    got = ob.point_cloud(depth, conv, ref, *cam, T)
    want = _numpy_points(depth, conv, ref, *cam, T)
    assert got.shape == want.shape == (int((conv == 1).sum()), 4)
    assert np.array_equal(got, want)


def test_oracle_point_cloud_known_answers_and_properties():
    # identity pose, principal point on a pixel: that pixel's ray is the optical axis
    depth = np.full((3, 5), 2.0, F)
    conv = np.ones((3, 5), np.int32)
    ref = np.arange(15, dtype=np.uint8).reshape(3, 5) * 10
    T = np.eye(4, dtype=F)[:3]
    pts = ob.point_cloud(depth, conv, ref, 100.0, 100.0, 2.0, 1.0, T)
    assert pts.shape == (15, 4)
    assert np.array_equal(pts[1 * 5 + 2], np.array([0, 0, 2.0, 70], F))        # pixel (x=2, y=1)
    assert np.array_equal(pts[:, 3], ref.ravel().astype(F))                        # row-major order, intensities
    assert np.allclose(np.linalg.norm(pts[:, :3], axis=1), 2.0, rtol=1e-6)         # |f| = 1: range == depth
    # only CONVERGED (== 1) pixels, whatever the other states are; translation is added last
    conv2 = np.array([[0, 1, 2, 3, 4], [5, 1, 1, 0, 0], [1, 0, 0, 0, 1]], np.int32)
    T2 = T.copy()
    T2[:, 3] = (10, 20, 30)
    pts2 = ob.point_cloud(depth, conv2, ref, 100.0, 100.0, 2.0, 1.0, T2)
    assert np.array_equal(pts2[:, 3], ref[conv2 == 1].astype(F)) and len(pts2) == 5
    assert np.allclose(np.linalg.norm(pts2[:, :3] - np.array([10, 20, 30], F), axis=1), 2.0, rtol=1e-6)
    assert len(ob.point_cloud(depth, np.zeros_like(conv), ref, 100.0, 100.0, 2.0, 1.0, T)) == 0   # empty cloud


# ------------------------------------------------------------------ GPU: the product
@pytest.mark.gpu
@pytest.mark.parametrize("size,frames", [((333, 201), 30), ((640, 480), 25), ((101, 77), 60)])
def test_product_point_cloud_equals_oracle(size, frames):
    import rpg_open_remode_b200 as rmd
    from rpg_open_remode_b200 import synth
    w, h = size
    seq = synth.SyntheticSequence(w, h, seed=0x5EED0300 + w)
    g = rmd.SeedMatrix(w, h, rmd.PinholeCamera(*seq.camera))
    f0 = seq.frame(0)
    dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
    g.setReferenceImage(f0.image_u8, f0.T_cam_world, dmin, dmax)
    for k in range(1, frames + 1):
        f = seq.frame(k, want_depth=False)
        g.update(f.image_u8, f.T_cam_world)
    cam = [float(F(c)) for c in seq.camera]
    T_world_ref = ob.se3_inv(f0.T_cam_world)
    mu, conv = g.downloadDepthmap(), g.downloadConvergence()
    want = ob.point_cloud(mu, conv, f0.image_u8, *cam, T_world_ref)
    assert len(want) > 0.05 * w * h, "the sequence is too short: hardly anything converged"
    pts, n = g.pointCloud()
    assert n == len(want) == g.getConvergedCount()
    assert np.array_equal(pts, want)
    # from a device depth image: the TV-L1 denoised map, as the node publishes it (src/depthmap_node.cpp denoiseAndPublishResults)
    den = rmd.DepthmapDenoiser(w, h)
    den.setLargeSigmaSq(dmax - dmin)
    dimg = rmd.DeviceImage(w, h, "float32")
    den.denoiseSeedsToDevice(g, dimg.data, dimg.pitch, 0.5, 30)
    den.sync()
    denoised = dimg.getDevData()
    pts_d, n_d = g.pointCloud(depth=dimg)
    assert n_d == n and np.array_equal(pts_d, ob.point_cloud(denoised, conv, f0.image_u8, *cam, T_world_ref))
    # capacity contract: count is always the full number, only `capacity` points are written, in order
    part, n_part = g.pointCloud(capacity=100)
    assert n_part == n and np.array_equal(part, want[:100])
    none, n_none = g.pointCloud(capacity=0)
    assert n_none == n and len(none) == 0
    # idempotent
    again, _ = g.pointCloud()
    assert np.array_equal(again, want)


@pytest.mark.gpu
def test_point_cloud_requires_reference_and_handles_empty(small_sequence):
    import rpg_open_remode_b200 as rmd
    seq = small_sequence
    g = rmd.SeedMatrix(seq.width, seq.height, rmd.PinholeCamera(*seq.camera))
    with pytest.raises(rmd.RmdError):
        g.pointCloud()
    f0 = seq.frame(0)
    g.setReferenceImage(f0.image, f0.T_cam_world, float(f0.depth.min()), float(f0.depth.max()))
    pts, n = g.pointCloud()          # nothing has converged right after the keyframe
    assert n == 0 and pts.shape == (0, 4)
