"""Pins of the CPU oracle (oracle/rmd_oracle.c) against the reference's own
known-answer tests, re-hosted on synthetic frames because the reference's data
set is external.  CPU only.

Each test names the reference gtest it re-hosts.  The GPU-side pin of the
oracle against the reference's rebuilt CUDA kernels is in
test_ref_cuda_parity.py.
"""
import numpy as np
import pytest

import oracle_binding as ob
from rpg_open_remode_b200 import synth

P = 5


def _ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


@pytest.fixture(scope="module")
def frames(small_sequence):
    return {k: small_sequence.frame(k) for k in (0, 1, 20)}


def _new_oracle(seq, patch=P):
    return ob.OracleSeeds(seq.width, seq.height, *seq.camera, patch=patch)


def test_seed_matrix_init(small_sequence, frames):
    """test/seed_matrix_test.cpp:29-151 (seedMatrixInit)."""
    seq, f = small_sequence, frames[1]
    min_d, max_d = 0.4, 1.8  # :66-67
    o = _new_oracle(seq)
    o.set_reference(f.image, f.T_cam_world, min_d, max_d)
    avg = np.float32((np.float32(min_d) + np.float32(max_d)) / np.float32(2))
    sig = np.float32((np.float32(max_d) - np.float32(min_d)) ** 2 / np.float32(36))
    # ASSERT_FLOAT_EQ == within 4 ulp, :99-110
    assert _ulp_diff(o.mu, avg).max() <= 4
    assert _ulp_diff(o.sigma_sq, sig).max() <= 4
    assert np.all(o.a == 10.0) and np.all(o.b == 10.0)
    # host computation in double over the interior, :120-150
    img = f.image.astype(np.float64)
    h, w = img.shape
    win = np.lib.stride_tricks.sliding_window_view(img, (P, P))  # [y, x] = patch with top-left (x, y)
    s = win.sum(axis=(2, 3))
    s2 = (win ** 2).sum(axis=(2, 3))
    ys, xs = slice(P, h - P // 2), slice(P, w - P // 2)
    exp_sum = s[P - P // 2:h - P // 2 - P // 2, P - P // 2:w - P // 2 - P // 2].astype(np.float32)
    exp_den = (P * P * s2 - s * s)[P - P // 2:h - P // 2 - P // 2, P - P // 2:w - P // 2 - P // 2].astype(np.float32)
    assert np.abs(o.sum_templ[ys, xs] - exp_sum).max() <= 1e-5 * 2  # fp32 sum of 25 terms vs double
    assert np.abs(o.const_templ_denom[ys, xs] - exp_den).max() <= 1e-3


def test_seed_matrix_check(small_sequence, frames):
    """test/seed_matrix_test.cpp:154-243 (seedMatrixCheck): the reference image
    is passed as the current image with the pose of frame 20 (:208-210)."""
    seq = small_sequence
    o = _new_oracle(seq)
    o.set_reference(frames[1].image, frames[1].T_cam_world, 0.4, 1.8)
    o.update(frames[1].image, frames[20].T_cam_world)
    conv = o.convergence
    h, w = conv.shape
    ring = np.ones((h, w), bool)
    ring[P:h - P, P:w - P] = False  # r > rows-P-1 || r < P || ... (:224-227)
    assert np.all(conv[ring] == ob.BORDER)
    allowed = [ob.UPDATE, ob.DIVERGED, ob.CONVERGED, ob.NOT_VISIBLE, ob.NO_MATCH]
    assert np.all(np.isin(conv[~ring], allowed))
    assert not np.any(conv[~ring] == ob.BORDER)


def test_epipolar_match_identity(small_sequence, frames):
    """test/epipolar_test.cpp:138-225 (epipolarMatchTest): reference frame used
    as current frame with its own pose; every pixel left in UPDATE must match
    itself within 0.01 px."""
    seq, f = small_sequence, frames[1]
    o = _new_oracle(seq)
    o.set_reference(f.image, f.T_cam_world, 0.4, 1.8)
    o.update(f.image, f.T_cam_world)
    conv, m = o.convergence, o.matches
    upd = conv == ob.UPDATE
    assert upd.sum() > 0.5 * (seq.width - 2 * P) * (seq.height - 2 * P), "self-match should succeed almost everywhere"
    ys, xs = np.nonzero(upd)
    assert np.abs(m[ys, xs, 0] - xs).max() <= 0.01
    assert np.abs(m[ys, xs, 1] - ys).max() <= 0.01


def test_reduction_sum_and_count():
    """test/reduction_test.cpp:24-122: 752x480 uniform [0,1) sum within 4 ulp of
    the double-accumulated sum; count of value 2 among uniform ints exact."""
    rng = np.random.default_rng(12345)
    img = rng.random((480, 752), dtype=np.float32)
    ref_order = ob.sum_f32_ref_order(img)
    dbl = np.float32(ob.sum_f32_f64(img))
    assert abs(ob.sum_f32_f64(img) - img.astype(np.float64).sum()) < 1e-6
    assert _ulp_diff(ref_order, dbl) <= 4
    ints = rng.integers(0, 256, size=(480, 752), dtype=np.int32)
    assert ob.count_equal_i32(ints, 2) == int((ints == 2).sum())
    assert ob.sum_i32(ints) == int(ints.sum())


def test_se3_helpers():
    """include/rmd/se3.cuh: quaternion ctor, inverse, product."""
    q = np.array([0.9, 0.1, -0.3, 0.2]); q /= np.linalg.norm(q)
    T = ob.se3_from_quat(*q.astype(np.float32), 0.3, -0.2, 1.5)
    R = T[:, :3].astype(np.float64)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-6) and np.isclose(np.linalg.det(R), 1.0, atol=1e-6)
    I = ob.se3_mul(T, ob.se3_inv(T))
    assert np.allclose(I[:, :3], np.eye(3), atol=1e-6) and np.allclose(I[:, 3], 0, atol=1e-6)


def test_filter_converges_to_ground_truth(small_sequence):
    """Accuracy sanity check the reference lacks (its .depth files are only used
    for min/max, test/dataset_main.cpp:70-77): after 30 frames most seeds have
    converged and the converged depths agree with the ray-cast ground truth."""
    seq = small_sequence
    f0 = seq.frame(0)
    o = _new_oracle(seq)
    o.set_reference(f0.image, f0.T_cam_world, float(f0.depth.min()), float(f0.depth.max()))
    for k in range(1, 31):
        f = seq.frame(k, want_depth=False)
        o.update(f.image, f.T_cam_world)
    conv = o.convergence
    interior = (seq.width - 2 * P) * (seq.height - 2 * P)
    c = conv == ob.CONVERGED
    assert c.sum() > 0.5 * interior
    err = np.abs(o.mu - f0.depth)[c]
    rng_d = float(f0.depth.max() - f0.depth.min())
    assert np.median(err) < 0.02 * rng_d
    assert np.percentile(err, 90) < 0.08 * rng_d
    assert o.converged_count() == int(c.sum())


def test_denoiser_smooths_and_keeps_range(small_sequence):
    """No reference test covers the denoiser; pin the
    properties of src/depthmap_denoiser.cu:62-118: the result stays within the
    input range, zero iterations return mu, and total variation drops."""
    rng = np.random.default_rng(3)
    h, w = 60, 80
    yy, xx = np.mgrid[0:h, 0:w]
    clean = (1.0 + 0.5 * (xx > w // 2) + 0.002 * yy).astype(np.float32)
    mu = (clean + 0.05 * rng.standard_normal((h, w))).astype(np.float32)
    sigma_sq = np.full((h, w), 1e-3, np.float32)
    a = np.full((h, w), 20.0, np.float32)
    b = np.full((h, w), 5.0, np.float32)
    out0 = ob.denoise(mu, sigma_sq, a, b, 1.0, 0.5, 0)
    assert np.array_equal(out0, mu)
    out = ob.denoise(mu, sigma_sq, a, b, 1.0, 0.5, 200)
    assert out.min() >= mu.min() - 1e-5 and out.max() <= mu.max() + 1e-5

    def tv(u):
        return np.abs(np.diff(u, axis=0)).sum() + np.abs(np.diff(u, axis=1)).sum()
    assert tv(out) < 0.5 * tv(mu)
    assert np.abs(out - clean).mean() < np.abs(mu - clean).mean()


def test_patch7_runs(small_sequence):
    """BASELINE config 4 uses a 7x7 patch (RMD_CORR_PATCH_SIDE=7): border ring 7."""
    seq = small_sequence
    f0, f1 = seq.frame(0), seq.frame(3, want_depth=False)
    o = _new_oracle(seq, patch=7)
    o.set_reference(f0.image, f0.T_cam_world, float(f0.depth.min()), float(f0.depth.max()))
    o.update(f1.image, f1.T_cam_world)
    conv = o.convergence
    assert np.all(conv[:7, :] == ob.BORDER) and np.all(conv[:, -7:] == ob.BORDER)
    assert np.all(conv[7:-7, 7:-7] != ob.BORDER)
