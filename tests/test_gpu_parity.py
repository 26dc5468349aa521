"""Parity of the sm_100a kernels (through the C-ABI) against the CPU oracle.

The GPU path is compiled -use_fast_math like the reference (CMakeLists.txt:25);
the oracle is IEEE fp32.  The two differ most in tau, the one-pixel
triangulation uncertainty (src/triangulation.cu:53-68: z_plus - z cancels 3-4
digits, so approximate sin/div move tau by ~1e-3 relative), which feeds sigma^2.
Tolerances below were set from the spread MEASURED on a B200 between the oracle
and the reference's own CUDA kernels rebuilt for sm_100a (oracle/_ref):
tests/perf/gpu_diag.py, profiles/r01_parity_spread.md.  They are the same spread:
our kernels agree with the reference's CUDA build far more tightly (see
test_ref_cuda_parity.py), so what is bounded here is IEEE-vs-fast-math.

  * seed initialisation: mu, sigma^2, a, b bit-exact; sum_templ <= 1e-5 and
    const_templ_denom <= 1e-3 absolute (the reference's own test tolerances,
    test/seed_matrix_test.cpp:148-149);
  * one update from bit-identical state: convergence states identical for
    >= 99.9 % of the pixels; among those |d mu| <= 1e-3 * depth_range and
    a, b within 1e-3 relative for >= 99.9 %, sigma^2 within 1e-2 relative for
    >= 99 % (measured: mu p99.9 4e-6, a/b p99.9 2e-5, sigma^2 p99 1e-3..2e-3);
    best match within 1e-3 px for >= 99.5 % (measured 99.96 %);
  * 30-frame sequences (differences compound through the search interval
    mu +- 3 sigma but stay bounded, the filter is contractive): states
    identical >= 99 % (measured 99.7 %); median |d mu| <= 1e-3 * range and
    >= 98 % within 1e-2 * range (measured median 2e-4, p99 3.5e-3); sigma^2
    median relative <= 1e-2 and >= 97 % within 0.2; a, b >= 97 % within 0.05;
  * denoiser vs the deterministic Jacobi oracle: max-abs <= 1e-4 * range
    (measured 7e-6; the reference's own run-to-run spread is 1e-3..5e-2);
  * reductions: counts exact, float sum within 4 ulp of the double sum.
"""
import numpy as np
import pytest

import oracle_binding as ob
import rpg_open_remode_b200 as rmd
from rpg_open_remode_b200 import synth

pytestmark = pytest.mark.gpu

P = 5
VARIANTS = [pytest.param(rmd.VARIANT_DIRECT, id="direct"), pytest.param(rmd.VARIANT_STAGED, id="staged")]


def _ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def _pair(seq, patch=P, variant=rmd.VARIANT_DIRECT, matches=True):
    cam = rmd.PinholeCamera(*seq.camera)
    g = rmd.SeedMatrix(seq.width, seq.height, cam, patch_side=patch)
    g.setOption(rmd.OPT_KERNEL_VARIANT, variant)
    if matches:
        g.setOption(rmd.OPT_RECORD_MATCHES, 1)
    o = ob.OracleSeeds(seq.width, seq.height, *seq.camera, patch=patch)
    return g, o


def _set_reference(seq, g, o):
    f0 = seq.frame(0)
    dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
    g.setReferenceImage(f0.image, f0.T_cam_world, dmin, dmax)
    o.set_reference(f0.image, f0.T_cam_world, dmin, dmax)
    return f0, dmax - dmin


def _field_stats(g, o, sel):
    out = {}
    for name, got, want in (("mu", g.downloadDepthmap(), o.mu), ("sigma_sq", g.downloadSigmaSq(), o.sigma_sq),
                            ("a", g.downloadA(), o.a), ("b", g.downloadB(), o.b)):
        d = np.abs(got.astype(np.float64) - want)[sel]
        out[name] = (d, d / np.maximum(np.abs(want.astype(np.float64))[sel], 1e-12))
    return out


def _assert_single_frame_parity(g, o, depth_range):
    """One update from bit-identical state (IEEE oracle vs fast-math GPU)."""
    conv_g, conv_o = g.downloadConvergence(), o.convergence
    same = conv_g == conv_o
    assert same.mean() >= 0.999, f"state agreement {same.mean():.5f}"
    st = _field_stats(g, o, same & (conv_o != ob.BORDER))
    assert (st["mu"][0] <= 1e-3 * depth_range).mean() >= 0.999
    assert (st["a"][1] <= 1e-3).mean() >= 0.999 and (st["b"][1] <= 1e-3).mean() >= 0.999
    assert (st["sigma_sq"][1] <= 1e-2).mean() >= 0.99, f"sigma_sq: {(st['sigma_sq'][1] <= 1e-2).mean():.5f}"


def _assert_sequence_parity(g, o, depth_range):
    """After a sequence (differences compound through mu +- 3 sigma)."""
    conv_g, conv_o = g.downloadConvergence(), o.convergence
    same = conv_g == conv_o
    assert same.mean() >= 0.99, f"state agreement {same.mean():.5f}"
    st = _field_stats(g, o, same & (conv_o != ob.BORDER))
    assert np.median(st["mu"][0]) <= 1e-3 * depth_range
    assert (st["mu"][0] <= 1e-2 * depth_range).mean() >= 0.98, f"mu: {(st['mu'][0] <= 1e-2 * depth_range).mean():.5f}"
    assert np.median(st["sigma_sq"][1]) <= 1e-2
    assert (st["sigma_sq"][1] <= 0.2).mean() >= 0.97
    assert (st["a"][1] <= 0.05).mean() >= 0.97 and (st["b"][1] <= 0.05).mean() >= 0.97


@pytest.mark.parametrize("patch", [5, 7])
def test_seed_init_parity(small_sequence, patch):
    """seedMatrixInit re-hosted (test/seed_matrix_test.cpp:99-150) on the GPU path."""
    g, o = _pair(small_sequence, patch)
    _set_reference(small_sequence, g, o)
    assert np.array_equal(g.downloadDepthmap(), o.mu)
    assert np.array_equal(g.downloadSigmaSq(), o.sigma_sq)
    assert np.array_equal(g.downloadA(), o.a) and np.array_equal(g.downloadB(), o.b)
    assert np.abs(g.downloadSumTempl() - o.sum_templ).max() <= 1e-5
    assert np.abs(g.downloadConstTemplDenom() - o.const_templ_denom).max() <= 1e-3
    conv = g.downloadConvergence()
    ring = np.ones_like(conv, bool)
    ring[patch:-patch, patch:-patch] = False
    assert np.all(conv[ring] == rmd.ConvergenceStates.BORDER)
    assert np.all(conv[~ring] == rmd.ConvergenceStates.UPDATE)
    assert g.getConvergedCount() == 0


@pytest.mark.parametrize("variant", VARIANTS)
def test_first_update_parity(small_sequence, variant):
    seq = small_sequence
    g, o = _pair(seq, variant=variant)
    _, rng_d = _set_reference(seq, g, o)
    f = seq.frame(4, want_depth=False)
    g.update(f.image, f.T_cam_world)
    o.update(f.image, f.T_cam_world)
    _assert_single_frame_parity(g, o, rng_d)
    both = (g.downloadConvergence() == 0) & (o.convergence == 0)
    dm = np.abs(g.downloadEpipolarMatches() - o.matches).max(axis=2)[both]
    assert (dm <= 1e-3).mean() >= 0.995
    assert abs(g.getDistFromRef() - o.dist_from_ref()) < 1e-6


@pytest.mark.parametrize("variant", VARIANTS)
def test_seed_matrix_check_and_identity_match(small_sequence, variant):
    """seedMatrixCheck (test/seed_matrix_test.cpp:154-243) and epipolarMatchTest
    (test/epipolar_test.cpp:138-225) re-hosted on the GPU path."""
    seq = small_sequence
    f1, f20 = seq.frame(1), seq.frame(20, want_depth=False)
    cam = rmd.PinholeCamera(*seq.camera)
    g = rmd.SeedMatrix(seq.width, seq.height, cam)
    g.setOption(rmd.OPT_KERNEL_VARIANT, variant)
    g.setOption(rmd.OPT_RECORD_MATCHES, 1)
    g.setReferenceImage(f1.image, rmd.SE3(f1.T_world_cam).inv(), 0.4, 1.8)
    g.update(f1.image, rmd.SE3(f20.T_world_cam).inv())
    conv = g.downloadConvergence()
    ring = np.ones_like(conv, bool)
    ring[P:-P, P:-P] = False
    S = rmd.ConvergenceStates
    assert np.all(conv[ring] == S.BORDER)
    assert np.all(np.isin(conv[~ring], [S.UPDATE, S.DIVERGED, S.CONVERGED, S.NOT_VISIBLE, S.NO_MATCH]))

    g2 = rmd.SeedMatrix(seq.width, seq.height, cam)
    g2.setOption(rmd.OPT_KERNEL_VARIANT, variant)
    g2.setOption(rmd.OPT_RECORD_MATCHES, 1)
    T = rmd.SE3(f1.T_world_cam).inv()
    g2.setReferenceImage(f1.image, T, 0.4, 1.8)
    g2.update(f1.image, T)
    conv, m = g2.downloadConvergence(), g2.downloadEpipolarMatches()
    ys, xs = np.nonzero(conv == S.UPDATE)
    assert len(ys) > 0.5 * (~ring).sum()
    assert np.abs(m[ys, xs, 0] - xs).max() <= 0.01 and np.abs(m[ys, xs, 1] - ys).max() <= 0.01


@pytest.mark.parametrize("variant", VARIANTS)
def test_update_from_uploaded_state(qvga_sequence, variant):
    """Kernel-level parity: both sides start one update from bit-identical state."""
    seq = qvga_sequence
    g, o = _pair(seq, variant=variant)
    _, rng_d = _set_reference(seq, g, o)
    for k in range(1, 13):
        f = seq.frame(k, want_depth=False)
        o.update(f.image, f.T_cam_world)
    for fid, arr in ((rmd.FIELD_MU, o.mu), (rmd.FIELD_SIGMA_SQ, o.sigma_sq), (rmd.FIELD_A, o.a),
                     (rmd.FIELD_B, o.b)):
        g.uploadState(fid, arr)
    assert np.array_equal(g.downloadDepthmap(), o.mu) and np.array_equal(g.downloadB(), o.b)
    f = seq.frame(13, want_depth=False)
    g.update(f.image, f.T_cam_world)
    o.update(f.image, f.T_cam_world)
    _assert_single_frame_parity(g, o, rng_d)
    assert g.getConvergedCount() == int((g.downloadConvergence() == 1).sum())


@pytest.mark.parametrize("variant", VARIANTS)
def test_sequence_parity_30_frames(qvga_sequence, variant):
    """BASELINE config 1 stand-in (30 frames, one keyframe) at QVGA."""
    seq = qvga_sequence
    g, o = _pair(seq, variant=variant, matches=False)
    f0, rng_d = _set_reference(seq, g, o)
    for k in range(1, 31):
        f = seq.frame(k, want_depth=False)
        g.update(f.image, f.T_cam_world)
        o.update(f.image, f.T_cam_world)
    _assert_sequence_parity(g, o, rng_d)
    conv = g.downloadConvergence()
    c = conv == 1
    assert c.sum() > 0.5 * (seq.width - 2 * P) * (seq.height - 2 * P)
    err = np.abs(g.downloadDepthmap() - f0.depth)[c]
    assert np.median(err) < 0.02 * rng_d
    assert g.getConvergedCount() == int(c.sum())


def test_patch7_sequence(small_sequence):
    seq = small_sequence
    g, o = _pair(seq, patch=7)
    _, rng_d = _set_reference(seq, g, o)
    for k in range(1, 6):
        f = seq.frame(k, want_depth=False)
        g.update(f.image, f.T_cam_world)
        o.update(f.image, f.T_cam_world)
    _assert_sequence_parity(g, o, rng_d)


def test_u8_and_depthmap_facade(small_sequence):
    """rmd::Depthmap call sequence (test/dataset_main.cpp:87-116): u8 frames in,
    conversion on the GPU == float frames converted on the host."""
    seq = small_sequence
    fx, fy, cx, cy = seq.camera
    f0 = seq.frame(0)
    dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
    dm = rmd.Depthmap(seq.width, seq.height, fx, cx, fy, cy)
    sm = rmd.SeedMatrix(seq.width, seq.height, rmd.PinholeCamera(fx, fy, cx, cy))
    assert dm.setReferenceImage(f0.image_u8, rmd.SE3(f0.T_world_cam).inv(), dmin, dmax)
    sm.setReferenceImage(f0.image, rmd.SE3(f0.T_world_cam).inv(), dmin, dmax)
    for k in range(1, 8):
        f = seq.frame(k, want_depth=False)
        T = rmd.SE3(f.T_world_cam).inv()
        dm.update(f.image_u8, T)
        sm.update(f.image, T)
    dm.downloadDepthmap()
    assert np.array_equal(dm.getDepthmap(), sm.downloadDepthmap())
    dm.downloadConvergenceMap()
    assert np.array_equal(dm.getConvergenceMap(), sm.downloadConvergence())
    assert dm.getConvergedCount() == sm.getConvergedCount()
    assert abs(dm.getConvergedPercentage() - 100.0 * dm.getConvergedCount() / (seq.width * seq.height)) < 1e-4
    assert dm.getDistFromRef() == sm.getDistFromRef() > 0
    dm.downloadDenoisedDepthmap(0.5, 20)
    assert dm.getDepthmap().shape == (seq.height, seq.width) and np.isfinite(dm.getDepthmap()).all()
    assert np.array_equal(dm.getReferenceImage(), f0.image_u8)


def test_absorbing_states_and_determinism(qvga_sequence):
    """Size-independent properties: CONVERGED / DIVERGED / BORDER seeds are never
    touched again (src/seed_update.cu:54-56) and two runs are bit-identical."""
    seq = qvga_sequence
    cam = rmd.PinholeCamera(*seq.camera)
    runs = []
    for _ in range(2):
        g = rmd.SeedMatrix(seq.width, seq.height, cam)
        f0 = seq.frame(0)
        g.setReferenceImage(f0.image, f0.T_cam_world, float(f0.depth.min()), float(f0.depth.max()))
        snaps = []
        for k in range(1, 31):
            f = seq.frame(k, want_depth=False)
            g.update(f.image, f.T_cam_world)
            if k in (24, 30):
                snaps.append((g.downloadConvergence(), g.downloadDepthmap(), g.downloadSigmaSq(),
                              g.downloadA(), g.downloadB()))
        runs.append(snaps)
    (c24, mu24, s24, a24, b24), (c30, mu30, s30, a30, b30) = runs[0]
    done = np.isin(c24, [1, 2, 3])
    assert done.sum() > 0
    assert np.array_equal(c24[done], c30[done])
    for x, y in ((mu24, mu30), (s24, s30), (a24, a30), (b24, b30)):
        assert np.array_equal(x[done], y[done])
    for s0, s1 in zip(runs[0], runs[1]):
        for x, y in zip(s0, s1):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("iters", [50, 200])
def test_denoiser_parity(qvga_sequence, iters):
    seq = qvga_sequence
    g, o = _pair(seq, matches=False)
    _, rng_d = _set_reference(seq, g, o)
    for k in range(1, 16):
        f = seq.frame(k, want_depth=False)
        g.update(f.image, f.T_cam_world)
    mu, s2, a, b = g.downloadDepthmap(), g.downloadSigmaSq(), g.downloadA(), g.downloadB()
    want = ob.denoise(mu, s2, a, b, rng_d, 0.5, iters)
    den = rmd.DepthmapDenoiser(seq.width, seq.height)
    den.setLargeSigmaSq(rng_d)
    got_seeds = den.denoiseSeeds(g, 0.5, iters)
    got_planar = den.denoise(g.getMu(), g.getSigmaSq(), g.getA(), g.getB(), 0.5, iters)
    assert np.array_equal(got_seeds, got_planar)
    assert np.abs(got_seeds - want).max() <= 1e-4 * rng_d
    # temporal blocking: 8 iterations per launch (csrc/denoiser.cu), plus the set-up kernel, two calls
    assert den.launchCount() == 2 * (1 + (iters + 7) // 8)


@pytest.mark.parametrize("size", [(101, 77), (33, 17), (320, 240), (49, 25)])
def test_denoiser_ragged_sizes_and_iteration_counts(size):
    """Tile seams (48 x 24 tiles + 8-pixel halo), partial tiles, odd widths (pixel pairs), images smaller than one
    tile, and iteration counts around the 8-iterations-per-launch block: all against the Jacobi oracle."""
    W, H = size
    seq = synth.SyntheticSequence(W, H, seed=0x5EED0020 + W)
    g, o = _pair(seq, matches=False)
    _, rng_d = _set_reference(seq, g, o)
    for k in range(1, 10):
        f = seq.frame(k, want_depth=False)
        g.update(f.image, f.T_cam_world)
    mu, s2, a, b = g.downloadDepthmap(), g.downloadSigmaSq(), g.downloadA(), g.downloadB()
    den = rmd.DepthmapDenoiser(W, H)
    den.setLargeSigmaSq(rng_d)
    for iters in (0, 1, 7, 8, 9, 16, 17, 40):
        want = ob.denoise(mu, s2, a, b, rng_d, 0.5, iters)
        got = den.denoiseSeeds(g, 0.5, iters)
        assert np.abs(got - want).max() <= 1e-4 * rng_d, (size, iters, float(np.abs(got - want).max() / rng_d))
    assert np.array_equal(den.denoiseSeeds(g, 0.5, 0), mu)


def test_denoiser_requires_large_sigma(small_sequence):
    """src/depthmap_denoiser.cu:189-193: denoise before setLargeSigmaSq is refused."""
    seq = small_sequence
    g, o = _pair(seq)
    _set_reference(seq, g, o)
    den = rmd.DepthmapDenoiser(seq.width, seq.height)
    with pytest.raises(rmd.RmdError):
        den.denoiseSeeds(g, 0.5, 10)


def test_update_requires_reference(small_sequence):
    seq = small_sequence
    g = rmd.SeedMatrix(seq.width, seq.height, rmd.PinholeCamera(*seq.camera))
    f = seq.frame(1, want_depth=False)
    with pytest.raises(rmd.RmdError):
        g.update(f.image, f.T_cam_world)


def test_reductions():
    """test/reduction_test.cpp:24-122 on the GPU path (752x480)."""
    rng = np.random.default_rng(12345)
    img = rng.random((480, 752), dtype=np.float32)
    d = rmd.DeviceImage(752, 480, "float32")
    d.setDevData(img)
    got = rmd.ImageReducer("float32").sum(d)
    assert _ulp_diff(got, np.float32(img.astype(np.float64).sum())) <= 4
    lo, hi = rmd.ImageReducer("float32").minMax(d)
    assert lo == img.min() and hi == img.max()
    ints = rng.integers(0, 256, size=(480, 752), dtype=np.int32)
    di = rmd.DeviceImage(752, 480, "int32")
    di.setDevData(ints)
    red = rmd.ImageReducer("int32")
    assert red.countEqual(di, 2) == int((ints == 2).sum()) == ob.count_equal_i32(ints, 2)
    assert red.sum(di) == int(ints.sum())
    # ragged sizes: a single row, a single column, a non-multiple-of-32 width
    for (w, h) in ((1, 1), (1, 37), (33, 1), (97, 13)):
        small = rng.integers(0, 3, size=(h, w), dtype=np.int32)
        ds = rmd.DeviceImage(w, h, "int32")
        ds.setDevData(small)
        assert red.countEqual(ds, 1) == int((small == 1).sum())
        assert red.sum(ds) == int(small.sum())


def test_device_image_round_trips():
    """test/device_image_test.cpp:27-115 (upload/download float, float2) + zero, copy."""
    rng = np.random.default_rng(1)
    a = rng.random((48, 70), dtype=np.float32)
    d = rmd.DeviceImage(70, 48, "float32")
    d.setDevData(a)
    assert np.array_equal(d.getDevData(), a)
    assert d.pitch >= 70 * 4 and d.stride * 4 == d.pitch
    e = rmd.DeviceImage(70, 48, "float32")
    e.assign(d)
    assert np.array_equal(e.getDevData(), a)
    e.zero()
    assert not e.getDevData().any()
    a2 = rng.random((48, 70, 2), dtype=np.float32)
    d2 = rmd.DeviceImage(70, 48, "float2")
    d2.setDevData(a2)
    assert np.array_equal(d2.getDevData(), a2)
