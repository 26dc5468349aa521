"""Parity at the lengths and on the paths that bench.py claims ("Next round" item 1).

(a) the device-resident entry points that `value` is timed through --
    setReferenceImageDevice / updateDevice / updateDeviceBatch on dense-pitch
    frames (TMA descriptors encoded on the caller's memory, not on the
    library's cudaMallocPitch ring) -- give the host path's result bit for bit;
(b) product vs the reference's OWN CUDA kernels (oracle/_ref, the unmodified
    /root/reference/src/*.cu rebuilt for sm_100a) over the FULL sequences of
    BASELINE configs 2, 3 and 4: protocol of test/dataset_main.cpp:87-116
    (frame 0 = reference view, every other frame one update, then the TV-L1
    denoiser), host sequencing of src/seed_matrix.cu:120-158.

Differences compound through mu +- 3 sigma (the next frame's search interval),
so the 30-frame bars of test_ref_cuda_parity.py do not carry over; the bars
here are set from the spread measured on a B200 and the measured agreement
is printed and written to gpurun_out/parity_full_length.json (copied to
profiles/r02_parity_full_length.json).
"""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
import ref_binding as rb
import rpg_open_remode_b200 as rmd
from rpg_open_remode_b200 import multi_gpu, synth

pytestmark = pytest.mark.gpu

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REPORT = os.path.join(_ROOT, "gpurun_out", "parity_full_length.json")


def _snap(g):
    return {"conv": g.downloadConvergence(), "mu": g.downloadDepthmap(), "sigma_sq": g.downloadSigmaSq(),
            "a": g.downloadA(), "b": g.downloadB()}


def _snap_ref(r):
    return {"conv": r.download(4), "mu": r.download(0), "sigma_sq": r.download(1), "a": r.download(2),
            "b": r.download(3)}


def agreement(A, B, depth_range):
    """Measured agreement of two seed states (A: product, B: reference)."""
    same = A["conv"] == B["conv"]
    sel = same & (B["conv"] != rmd.ConvergenceStates.BORDER)
    d = np.abs(A["mu"].astype(np.float64) - B["mu"])[sel] / depth_range
    out = {"pixels": int(same.size), "states_equal": float(same.mean()),
           "mu_bit_identical": float((d == 0).mean()), "mu_within_1e-3_range": float((d <= 1e-3).mean()),
           "mu_p99_over_range": float(np.percentile(d, 99)), "mu_max_over_range": float(d.max()),
           "converged_product": int((A["conv"] == 1).sum()), "converged_reference": int((B["conv"] == 1).sum())}
    # Seeds whose posterior variance has collapsed to a rounding residue of either sign (sigma_sq <= 0 in
    # either implementation) are "trapped": sqrtf(sigma_sq) is NaN, both report NO_MATCH and add 1 to b every
    # frame (defined deviation 2).  WHEN a seed falls in depends on the last bit of its variance,
    # so their b counts are compared as a population, not pixel by pixel.
    trapped_a, trapped_b = A["sigma_sq"] <= 0, B["sigma_sq"] <= 0
    out["trapped_product"], out["trapped_reference"] = int(trapped_a.sum()), int(trapped_b.sum())
    live = sel & ~trapped_a & ~trapped_b
    for name, tol in (("sigma_sq", 1e-2), ("a", 1e-3), ("b", 1e-3)):
        rel = (np.abs(A[name].astype(np.float64) - B[name]) / np.maximum(np.abs(B[name]), 1e-12))[live]
        out[f"{name}_within_{tol:g}_rel"] = float((rel <= tol).mean())
    return out


def _report(key, value):
    os.makedirs(os.path.dirname(_REPORT), exist_ok=True)
    try:
        with open(_REPORT) as f:
            rep = json.load(f)
    except Exception:
        rep = {}
    rep[key] = value
    with open(_REPORT, "w") as f:
        json.dump(rep, f, indent=1, sort_keys=True)
    print(f"[parity] {key}: " + json.dumps(value))


def _render(seq, n):
    frames = [seq.frame(k, want_depth=(k == 0)) for k in range(n)]
    f0 = frames[0]
    return frames, float(f0.depth.min()), float(f0.depth.max())


# ------------------------------------------------------------------ (a)

@pytest.mark.parametrize("patch,size", [(5, (640, 480)), (7, (352, 264))])
def test_device_resident_path_equals_host_path(patch, size):
    """What bench.py times as `value` (setReferenceImageDevice + updateDeviceBatch on dense frames) and the
    per-launch pass (updateDevice) == the host path of every other test, bit for bit."""
    import torch
    W, H = size
    N = 48
    seq = synth.SyntheticSequence(W, H, seed=multi_gpu.keyframe_seed(0))
    frames, dmin, dmax = _render(seq, N)
    poses = np.stack([f.T_cam_world.reshape(12) for f in frames]).astype(np.float32)
    cam = rmd.PinholeCamera(*seq.camera)
    dev = torch.device("cuda", 0)
    dense = torch.from_numpy(np.stack([f.image for f in frames])).to(dev)      # pitch = W * 4: dense rows
    assert dense.stride(1) == W and dense.data_ptr() % 16 == 0

    host = rmd.SeedMatrix(W, H, cam, patch_side=patch)
    host.setReferenceImage(frames[0].image, poses[0], dmin, dmax)
    for k in range(1, N):
        host.update(frames[k].image, poses[k])
    want = _snap(host)

    batch = rmd.SeedMatrix(W, H, cam, patch_side=patch)
    batch.setOption(rmd.OPT_CHAIN_FRAMES, 8)
    batch.setReferenceImageDevice(dense[0].data_ptr(), W * 4, poses[0], dmin, dmax)
    batch.updateDeviceBatch(dense[1].data_ptr(), W * H * 4, W * 4, poses[1:])
    single = rmd.SeedMatrix(W, H, cam, patch_side=patch)
    single.setReferenceImageDevice(dense[0].data_ptr(), W * 4, poses[0], dmin, dmax)
    for k in range(1, N):
        single.updateDevice(dense[k].data_ptr(), W * 4, poses[k])
    # the same with 3 frames and with 1 frame (the default) per launch (RMD_OPT_CHAIN_FRAMES)
    others = []
    for frames_per_launch in (3, 1):
        g = rmd.SeedMatrix(W, H, cam, patch_side=patch)
        g.setOption(rmd.OPT_CHAIN_FRAMES, frames_per_launch)
        g.setReferenceImageDevice(dense[0].data_ptr(), W * 4, poses[0], dmin, dmax)
        g.updateDeviceBatch(dense[1].data_ptr(), W * H * 4, W * 4, poses[1:20])      # two calls: chains restart cleanly
        g.updateDeviceBatch(dense[20].data_ptr(), W * H * 4, W * 4, poses[20:])
        others.append(("chain of %d" % frames_per_launch, g))
    for name, g in [("updateDeviceBatch", batch), ("updateDevice", single)] + others:
        got = _snap(g)
        for field in ("conv", "mu", "sigma_sq", "a", "b"):
            assert np.array_equal(got[field], want[field]), f"{name}: {field} differs from the host path"
        assert g.getConvergedCount() == host.getConvergedCount()
        g.sync()      # also reports an expired device-side wait of a chained launch
    assert batch.launchCount()[0] == (N - 1 + 7) // 8 and single.launchCount()[0] == N - 1
    torch.cuda.synchronize()


# ------------------------------------------------------------------ (b)

def _run_full(seq, n_frames, patch, every=None):
    """Product and reference CUDA over the same n_frames-1 updates (host float frames through update())."""
    f0 = seq.frame(0)
    dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
    g = rmd.SeedMatrix(seq.width, seq.height, rmd.PinholeCamera(*seq.camera), patch_side=patch)
    r = rb.RefSeeds(seq.width, seq.height, *seq.camera, patch=patch)
    g.setReferenceImage(f0.image, f0.T_cam_world, dmin, dmax)
    r.set_reference(f0.image, f0.T_cam_world, dmin, dmax)
    trace = {}
    for k in range(1, n_frames):
        f = seq.frame(k, want_depth=False)
        g.update(f.image, f.T_cam_world)
        r.update(f.image, f.T_cam_world)
        if every and (k % every == 0):
            trace[k] = agreement(_snap(g), _snap_ref(r), dmax - dmin)
    return g, r, dmax - dmin, f0, trace


# Bars: (states equal, mu bit-identical, mu within 1e-3 range, sigma_sq/a/b within tolerance).
# Set from the measured agreement (profiles/r02_parity_full_length.json), with margin.
# Measured (profiles/r02_parity_full_length.json): states 99.82-99.97 %, mu bit-identical 97.4-98.1 %, mu within
# 1e-3 range 99.6-99.8 %, sigma_sq / a / b 97.8-99.6 %, converged counts equal to 3-50 of 0.3-2.1 M pixels.
BARS = {
    "c2": dict(states=0.998, identical=0.95, within=0.99, params=0.97),
    "c3": dict(states=0.995, identical=0.95, within=0.99, params=0.96),
    "c4": dict(states=0.995, identical=0.95, within=0.99, params=0.96),
}


def _assert_bars(m, bars):
    assert m["states_equal"] >= bars["states"], m
    assert m["mu_bit_identical"] >= bars["identical"], m
    assert m["mu_within_1e-3_range"] >= bars["within"], m
    for k in ("sigma_sq_within_0.01_rel", "a_within_0.001_rel", "b_within_0.001_rel"):
        assert m[k] >= bars["params"], (k, m)
    n = m["pixels"]
    assert abs(m["converged_product"] - m["converged_reference"]) <= 5e-3 * n, m
    assert abs(m["trapped_product"] - m["trapped_reference"]) <= 5e-3 * n, m


@pytest.mark.skipif(not rb.available(5), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("keyframe", [0, 1])
def test_c2_full_sequence_vs_reference_cuda(keyframe):
    """BASELINE configs[1], the bench workload itself: VGA, 199 updates, keyframe seeds 0 and 1."""
    seq = synth.SyntheticSequence(640, 480, seed=multi_gpu.keyframe_seed(keyframe))
    g, r, rng_d, f0, trace = _run_full(seq, 200, 5, every=50)
    m = agreement(_snap(g), _snap_ref(r), rng_d)
    m["trace_states_equal"] = {str(k): v["states_equal"] for k, v in trace.items()}
    m["trace_mu_bit_identical"] = {str(k): v["mu_bit_identical"] for k, v in trace.items()}
    _report(f"c2_vga_199_updates_keyframe{keyframe}", m)
    _assert_bars(m, BARS["c2"])
    c = g.downloadConvergence() == 1
    assert np.median(np.abs(g.downloadDepthmap() - f0.depth)[c]) < 0.01 * rng_d


@pytest.mark.skipif(not rb.available(5), reason="oracle/_ref not built")
def test_c3_full_sequence_and_denoiser_vs_reference_cuda():
    """BASELINE configs[2]: 1280x720, 499 updates, then TV-L1 with 50 and 200 iterations
    (test/dataset_main.cpp:116 uses 0.5 / 200; BASELINE asks for 50)."""
    W, H = 1280, 720
    seq = synth.SyntheticSequence(W, H, seed=0x5EED0003)
    g, r, rng_d, f0, trace = _run_full(seq, 500, 5, every=100)
    A, B = _snap(g), _snap_ref(r)
    m = agreement(A, B, rng_d)
    m["trace_states_equal"] = {str(k): v["states_equal"] for k, v in trace.items()}
    _report("c3_720p_499_updates", m)
    _assert_bars(m, BARS["c3"])
    # the denoiser on IDENTICAL input (the reference's final state), so that only the solver is compared
    for fid in (0, 1, 2, 3):
        g.uploadState(fid, r.download(fid))
    den = rmd.DepthmapDenoiser(W, H)
    den.setLargeSigmaSq(rng_d)
    rden = rb.RefDenoiser(W, H)
    s = _snap(g)
    for iters in (50, 200):
        mine = den.denoiseSeeds(g, 0.5, iters)
        r1 = rden.run(r, rng_d, 0.5, iters)
        r2 = rden.run(r, rng_d, 0.5, iters)
        jac = ob.denoise(s["mu"], s["sigma_sq"], s["a"], s["b"], rng_d, 0.5, iters)
        d_ref = np.abs(mine - r1) / rng_d
        d_jac = np.abs(mine - jac) / rng_d
        rep = {"vs_reference_median": float(np.median(d_ref)), "vs_reference_p99": float(np.percentile(d_ref, 99)),
               "vs_reference_max": float(d_ref.max()),
               "reference_run_to_run_max": float(np.abs(r1 - r2).max() / rng_d),
               "vs_jacobi_oracle_max": float(d_jac.max()),
               "vs_jacobi_oracle_within_1e-4": float((d_jac <= 1e-4).mean())}
        _report(f"c3_720p_denoise_{iters}_iterations", rep)
        # the reference is racy across its 16x16 tile seams and does not reproduce itself;
        # ours is the deterministic Jacobi limit: within 1e-4 range of the Jacobi oracle (a last-ulp
        # difference may flip one shrink branch, worth 2 tau lambda at that pixel)
        assert (d_jac <= 1e-4).mean() >= 0.9999 and d_jac.max() <= 2e-2
        assert np.median(d_ref) <= 1e-4 and np.percentile(d_ref, 99) <= 2e-2


@pytest.mark.skipif(not rb.available(7), reason="oracle/_ref/librmd_ref_p7.so not built")
def test_c4_full_sequence_vs_reference_cuda():
    """BASELINE configs[3]: 1920x1080, 7x7 NCC (RMD_CORR_PATCH_SIDE=7, CMakeLists.txt:51), 499 updates."""
    W, H = 1920, 1080
    seq = synth.SyntheticSequence(W, H, seed=0x5EED0004)
    g, r, rng_d, f0, trace = _run_full(seq, 500, 7, every=100)
    A, B = _snap(g), _snap_ref(r)
    m = agreement(A, B, rng_d)
    m["trace_states_equal"] = {str(k): v["states_equal"] for k, v in trace.items()}
    _report("c4_1080p_p7_499_updates", m)
    _assert_bars(m, BARS["c4"])
    ring = np.ones((H, W), bool)
    ring[7:-7, 7:-7] = False
    assert np.all(A["conv"][ring] == 2) and np.all(B["conv"][ring] == 2)
