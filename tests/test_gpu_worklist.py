"""The staged kernel's bookkeeping -- per-frame work list, tile retirement, tile splitting,
sparse-tile path, programmatic dependent launch -- must never show in the results.

Everything here compares bit for bit against the direct variant (one thread per pixel, no
bookkeeping at all; itself checked against the oracle and the reference's CUDA build in
test_gpu_parity.py / test_ref_cuda_parity.py), on cases chosen to stress the bookkeeping:
ragged image sizes (partial tiles), switching variants and uploading state in the middle of
a keyframe (the work list has to be rebuilt), extreme settings of every tuning knob, and a
long sequence in which most tiles retire (the converged count must keep counting them).
"""
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


def _new(seq, variant, patch=5, knobs=()):
    g = rmd.SeedMatrix(seq.width, seq.height, rmd.PinholeCamera(*seq.camera), patch_side=patch)
    g.setOption(rmd.OPT_KERNEL_VARIANT, variant)
    for opt, val in knobs:
        g.setOption(opt, val)
    f0 = seq.frame(0)
    g.setReferenceImage(f0.image, f0.T_cam_world, float(f0.depth.min()), float(f0.depth.max()))
    return g


def _run(g, seq, first, last):
    for k in range(first, last + 1):
        f = seq.frame(k, want_depth=False)
        g.update(f.image, f.T_cam_world)


@pytest.mark.parametrize("size,patch", [((101, 77), 5), ((70, 50), 5), ((203, 131), 7), ((33, 17), 5)])
def test_ragged_sizes_staged_equals_direct(size, patch):
    """Width not a multiple of 32, height not a multiple of 8, down to a single partial tile row."""
    seq = synth.SyntheticSequence(size[0], size[1], seed=0x5EED0010 + size[0])
    d, s = _new(seq, rmd.VARIANT_DIRECT, patch), _new(seq, rmd.VARIANT_STAGED, patch)
    for k in range(1, 16):
        f = seq.frame(k, want_depth=False)
        d.update(f.image, f.T_cam_world)
        s.update(f.image, f.T_cam_world)
        if k in (1, 7, 15):
            D, S = _snap(d), _snap(s)
            _same(S, D, f"{size} frame {k}")
            assert s.getConvergedCount() == d.getConvergedCount() == int((D["conv"] == 1).sum())
    ring = np.ones((size[1], size[0]), bool)
    ring[patch:-patch, patch:-patch] = False
    assert np.all(S["conv"][ring] == 2)


def test_variant_switch_and_state_upload_mid_keyframe(qvga_sequence):
    seq = qvga_sequence
    ref = _new(seq, rmd.VARIANT_DIRECT)
    _run(ref, seq, 1, 45)
    want = _snap(ref)

    # staged -> direct -> staged: the work list is rebuilt, retired tiles are counted again
    g = _new(seq, rmd.VARIANT_STAGED)
    _run(g, seq, 1, 20)
    c20 = g.getConvergedCount()
    assert c20 == int((g.downloadConvergence() == 1).sum())
    g.setOption(rmd.OPT_KERNEL_VARIANT, rmd.VARIANT_DIRECT)
    _run(g, seq, 21, 26)
    assert g.getConvergedCount() == int((g.downloadConvergence() == 1).sum())
    g.setOption(rmd.OPT_KERNEL_VARIANT, rmd.VARIANT_STAGED)
    _run(g, seq, 27, 45)
    _same(_snap(g), want, "staged/direct/staged")
    assert g.getConvergedCount() == int((want["conv"] == 1).sum())

    # the whole state moved into a fresh handle in the middle of the keyframe
    a = _new(seq, rmd.VARIANT_STAGED)
    _run(a, seq, 1, 20)
    st = _snap(a)
    b = _new(seq, rmd.VARIANT_STAGED)
    _run(b, seq, 1, 3)   # some other state first: uploads must fully replace it
    for field, name in ((rmd.FIELD_MU, "mu"), (rmd.FIELD_SIGMA_SQ, "sigma_sq"), (rmd.FIELD_A, "a"), (rmd.FIELD_B, "b"),
                        (rmd.FIELD_CONVERGENCE, "conv")):
        b.uploadState(field, st[name])
    _run(b, seq, 21, 45)
    _same(_snap(b), want, "uploaded state")
    assert b.getConvergedCount() == int((want["conv"] == 1).sum())


KNOB_SETS = {
    "no_split_no_sparse": [(rmd.OPT_TUNE_SPLIT_MAX, 1), (rmd.OPT_TUNE_SPARSE_MAX_SEEDS, 0), (rmd.OPT_TUNE_PDL, 0)],
    "split_everything": [(rmd.OPT_TUNE_SPLIT_MAX, 32), (rmd.OPT_TUNE_SPLIT_MIN_ITEMS, 1),
                         (rmd.OPT_TUNE_SPLIT_ITEMS_PER_CTA, 1), (rmd.OPT_TUNE_SPLIT_AVG_PCT, 1),
                         (rmd.OPT_TUNE_HEAVY_MIN_ITEMS, 1)],
    "all_sparse": [(rmd.OPT_TUNE_SPARSE_MAX_SEEDS, 256), (rmd.OPT_TUNE_SPLIT_MAX, 1), (rmd.OPT_TUNE_PDL, 2)],
    "nothing_heavy": [(rmd.OPT_TUNE_HEAVY_MIN_ITEMS, 65535), (rmd.OPT_TUNE_SPLIT_MAX, 4),
                      (rmd.OPT_TUNE_SPLIT_MIN_ITEMS, 64), (rmd.OPT_TUNE_SPLIT_ITEMS_PER_CTA, 32)],
}


@pytest.mark.parametrize("name", sorted(KNOB_SETS))
def test_tuning_knobs_do_not_change_results(qvga_sequence, name):
    seq = qvga_sequence
    ref = _new(seq, rmd.VARIANT_DIRECT)
    g = _new(seq, rmd.VARIANT_STAGED, knobs=KNOB_SETS[name])
    for k in range(1, 41):
        f = seq.frame(k, want_depth=False)
        ref.update(f.image, f.T_cam_world)
        g.update(f.image, f.T_cam_world)
        if k in (2, 15, 40):
            _same(_snap(g), _snap(ref), f"{name} frame {k}")
            assert g.getConvergedCount() == ref.getConvergedCount()


def test_retired_tiles_keep_counting(qvga_sequence):
    """After 120 frames most 32x8 tiles hold only absorbing seeds and have left the work list."""
    seq = qvga_sequence
    ref, g = _new(seq, rmd.VARIANT_DIRECT), _new(seq, rmd.VARIANT_STAGED)
    g.setOption(rmd.OPT_DEBUG_TIMELINE, 1)
    done_at_90 = None
    for k in range(1, 121):
        f = seq.frame(k, want_depth=False)
        ref.update(f.image, f.T_cam_world)
        g.update(f.image, f.T_cam_world)
        if k % 30 == 0:
            S = _snap(g)
            _same(S, _snap(ref), f"frame {k}")
            assert g.getConvergedCount() == ref.getConvergedCount() == int((S["conv"] == 1).sum())
        if k == 90:
            tiles = S["conv"].reshape(seq.height // 8, 8, seq.width // 32, 32)
            done_at_90 = int(np.isin(tiles, [1, 2, 3]).all(axis=(1, 3)).sum())
    # the debug timeline has a start stamp for every tile the last launch still visited
    listed = int((g.downloadTimeline()[:, 0] > 0).sum())
    n_tiles = ((seq.width + 31) // 32) * ((seq.height + 7) // 8)
    assert done_at_90 > 0, "the sequence is too short for this test: no tile was finished after 90 frames"
    assert 0 < listed <= n_tiles - done_at_90, f"{listed} of {n_tiles} tiles listed, {done_at_90} were finished 30 frames ago"
