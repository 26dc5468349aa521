"""GPU-box diagnostic: measures the spread between the three implementations
(ours / CPU oracle / the reference's rebuilt CUDA kernels) so the tolerances in
tests/ are set from data, and times both CUDA paths per frame.

    python tests/perf/gpu_diag.py [--width 320 --height 240 --frames 30] > gpurun_out/diag.txt
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_binding as ob  # noqa: E402
import ref_binding as rb  # noqa: E402
import rpg_open_remode_b200 as rmd  # noqa: E402
from rpg_open_remode_b200 import synth  # noqa: E402

FIELDS = {"mu": 0, "sigma_sq": 1, "a": 2, "b": 3}


def compare(name, A, B, rng_d):
    """A, B: dicts with conv, mu, sigma_sq, a, b (+ matches)."""
    out = {"pair": name}
    same = A["conv"] == B["conv"]
    out["state_agree"] = float(same.mean())
    interior = A["conv"] != 2
    out["state_agree_interior"] = float(same[interior].mean())
    for f in FIELDS:
        d = np.abs(A[f].astype(np.float64) - B[f].astype(np.float64))[interior & same]
        scale = rng_d if f == "mu" else np.maximum(np.abs(B[f].astype(np.float64))[interior & same], 1e-12)
        rel = d / scale
        fin = np.isfinite(rel)
        out[f] = {"median": float(np.median(rel[fin])), "p99": float(np.percentile(rel[fin], 99)),
                  "p999": float(np.percentile(rel[fin], 99.9)), "max": float(rel[fin].max()),
                  "frac_gt_1e-3": float((rel[fin] > 1e-3).mean()), "nonfinite": int((~fin).sum())}
    if "matches" in A and "matches" in B:
        both = (A["conv"] == 0) & (B["conv"] == 0)
        dm = np.abs(A["matches"] - B["matches"]).max(axis=2)[both]
        out["match"] = {"n": int(both.sum()), "frac_gt_1e-3px": float((dm > 1e-3).mean()),
                        "frac_gt_0.5px": float((dm > 0.5).mean()), "max": float(dm.max())}
    return out


def snapshot_ours(s, matches=True):
    d = {"conv": s.downloadConvergence(), "mu": s.downloadDepthmap(), "sigma_sq": s.downloadSigmaSq(),
         "a": s.downloadA(), "b": s.downloadB()}
    if matches:
        d["matches"] = s.downloadEpipolarMatches()
    return d


def snapshot_oracle(o):
    return {"conv": o.convergence.copy(), "mu": o.mu.copy(), "sigma_sq": o.sigma_sq.copy(), "a": o.a.copy(),
            "b": o.b.copy(), "matches": o.matches.copy()}


def snapshot_ref(r):
    return {"conv": r.download(4), "mu": r.download(0), "sigma_sq": r.download(1), "a": r.download(2),
            "b": r.download(3), "matches": r.download(7)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--height", type=int, default=240)
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--patch", type=int, default=5)
    ap.add_argument("--frac-bits", type=int, default=8)
    ap.add_argument("--no-oracle", action="store_true")
    args = ap.parse_args()
    W, H, N, PS = args.width, args.height, args.frames, args.patch

    seq = synth.SyntheticSequence(W, H, seed=0x5EED0001)
    frames = [seq.frame(k, want_depth=(k == 0)) for k in range(N + 1)]
    f0 = frames[0]
    dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
    rng_d = dmax - dmin
    cam = rmd.PinholeCamera(*seq.camera)

    ours = rmd.SeedMatrix(W, H, cam, patch_side=PS)
    ours.setOption(rmd.OPT_RECORD_MATCHES, 1)
    ours.setOption(rmd.OPT_KERNEL_VARIANT, rmd.VARIANT_DIRECT)
    ours.setOption(rmd.OPT_TEX_FRAC_BITS, args.frac_bits)
    ours.enableKernelTiming(True)
    ours.setReferenceImage(f0.image, f0.T_cam_world, dmin, dmax)
    ref = rb.RefSeeds(W, H, *seq.camera, patch=PS) if rb.available(PS) else None
    if ref:
        ref.set_reference(f0.image, f0.T_cam_world, dmin, dmax)
    orc = None
    if not args.no_oracle:
        orc = ob.OracleSeeds(W, H, *seq.camera, patch=PS, tex_frac_bits=args.frac_bits)
        orc.set_reference(f0.image, f0.T_cam_world, dmin, dmax)

    report = {"config": vars(args), "depth_range": rng_d, "frames": []}
    # init parity
    init = {"sum_templ_maxabs": None}
    if orc is not None:
        init["sum_templ_maxabs_vs_oracle"] = float(np.abs(ours.downloadSumTempl() - orc.sum_templ).max())
        init["denom_maxabs_vs_oracle"] = float(np.abs(ours.downloadConstTemplDenom() - orc.const_templ_denom).max())
        init["mu_equal"] = bool(np.array_equal(ours.downloadDepthmap(), orc.mu))
    if ref:
        init["sum_templ_maxabs_vs_ref"] = float(np.abs(ours.downloadSumTempl() - ref.download(5)).max())
        init["denom_maxabs_vs_ref"] = float(np.abs(ours.downloadConstTemplDenom() - ref.download(6)).max())
    report["init"] = init

    for k in range(1, N + 1):
        f = frames[k]
        T = f.T_cam_world
        ours.update(f.image, T)
        ours.sync()
        t_ours = ours.lastKernelMs()
        t_ref = None
        if ref:
            ref.sync()
            t0 = time.perf_counter()
            ref.update(f.image, T)
            ref.sync()
            t_ref = (time.perf_counter() - t0) * 1e3
        t_orc = None
        if orc is not None:
            t0 = time.perf_counter()
            orc.update(f.image, T)
            t_orc = (time.perf_counter() - t0) * 1e3
        entry = {"k": k, "ours_kernel_ms": t_ours, "ref_update_wall_ms": t_ref, "oracle_ms": t_orc,
                 "ours_converged": ours.getConvergedCount()}
        if k in (1, 2, 5, 10, 20, N):
            so = snapshot_ours(ours)
            if orc is not None:
                entry["ours_vs_oracle"] = compare("ours/oracle", so, snapshot_oracle(orc), rng_d)
            if ref:
                sr = snapshot_ref(ref)
                entry["ours_vs_ref"] = compare("ours/refcuda", so, sr, rng_d)
                if orc is not None:
                    entry["oracle_vs_ref"] = compare("oracle/refcuda", snapshot_oracle(orc), sr, rng_d)
            cnt = np.bincount(so["conv"].ravel(), minlength=6).tolist()
            entry["ours_state_hist"] = cnt
            gt_err = np.abs(so["mu"] - f0.depth)[so["conv"] == 1]
            entry["ours_gt_median_err"] = float(np.median(gt_err)) if gt_err.size else None
        report["frames"].append(entry)

    # single-frame parity from identical state: upload the oracle's state into ours and ref
    if orc is not None:
        f = seq.frame(N + 1, want_depth=False)
        st = snapshot_oracle(orc)
        for name, fid in FIELDS.items():
            ours.uploadState(fid, st[name])
            if ref:
                ref.upload(fid, st[name])
        ours.update(f.image, f.T_cam_world)
        orc.update(f.image, f.T_cam_world)
        so, sc = snapshot_ours(ours), snapshot_oracle(orc)
        report["one_frame_same_state"] = {"ours_vs_oracle": compare("ours/oracle", so, sc, rng_d)}
        if ref:
            ref.update(f.image, f.T_cam_world)
            sr = snapshot_ref(ref)
            report["one_frame_same_state"]["ours_vs_ref"] = compare("ours/refcuda", so, sr, rng_d)
            report["one_frame_same_state"]["oracle_vs_ref"] = compare("oracle/refcuda", sc, sr, rng_d)

    # denoiser
    den = rmd.DepthmapDenoiser(W, H)
    den.setLargeSigmaSq(rng_d)
    dres = {}
    for iters in (50, 200):
        mine = den.denoiseSeeds(ours, 0.5, iters)
        entry = {}
        if orc is not None:
            so = snapshot_ours(ours, matches=False)
            want = ob.denoise(so["mu"], so["sigma_sq"], so["a"], so["b"], rng_d, 0.5, iters)
            entry["ours_vs_oracle_maxabs_over_range"] = float(np.abs(mine - want).max() / rng_d)
        if ref:
            rden = rb.RefDenoiser(W, H, patch=PS)
            r1 = rden.run(ref, rng_d, 0.5, iters)
            r2 = rden.run(ref, rng_d, 0.5, iters)
            entry["ref_run_to_run_maxabs_over_range"] = float(np.abs(r1 - r2).max() / rng_d)
            if orc is not None:
                sr = snapshot_ref(ref)
                want_r = ob.denoise(sr["mu"], sr["sigma_sq"], sr["a"], sr["b"], rng_d, 0.5, iters)
                entry["ref_vs_oracle_jacobi_maxabs_over_range"] = float(np.abs(r1 - want_r).max() / rng_d)
        dres[str(iters)] = entry
    report["denoiser"] = dres
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
