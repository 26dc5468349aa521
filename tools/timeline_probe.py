"""Per-CTA timeline of the staged kernel at a steady-state and a heavy frame (GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpg_open_remode_b200 as rmd
from rpg_open_remode_b200 import synth

W, H = 640, 480
seq = synth.SyntheticSequence(W, H, seed=0x5EED0002)
f0 = seq.frame(0)
s = rmd.SeedMatrix(W, H, rmd.PinholeCamera(*seq.camera))
s.setOption(rmd.OPT_KERNEL_VARIANT, rmd.VARIANT_STAGED)
s.setOption(rmd.OPT_DEBUG_TIMELINE, 1)
for opt, val in zip((10, 11, 12, 13, 14, 15), [int(v) for v in sys.argv[1:7]]):
    s.setOption(opt, val)
s.enableKernelTiming(True)
s.setReferenceImage(f0.image, f0.T_cam_world, float(f0.depth.min()), float(f0.depth.max()))
for k in range(1, 131):
    f = seq.frame(k, want_depth=False)
    s.update(f.image, f.T_cam_world)
    if k in (4, 40, 75, 130):
        s.sync()
        t = s.downloadTimeline()
        n_all = len(t)
        ids = np.nonzero(t[:, 0] > 0)[0]
        t = t[ids]                  # tiles on this frame's work list (retired tiles are not launched)
        ms = s.lastKernelMs()
        act = t[:, 1] > 0
        start = t[:, 0] - t[:, 0].min()
        end = t[:, 5] - t[:, 0].min()
        dur_ns = (t[:, 5] - t[:, 0])[act]
        print(f"frame {k}: kernel {ms*1e3:.1f} us; tiles {n_all}, listed {len(t)}, active {act.sum()}, items total {t[:,7].sum()}, "
              f"span of starts {start.max()/1e3:.1f} us, last end {end[act].max()/1e3 if act.any() else 0:.1f} us")
        if act.any():
            ph = t[act][:, 1:5].astype(np.float64)
            d = np.diff(np.concatenate([np.zeros((ph.shape[0], 1)), ph], axis=1), axis=1)
            for name, col in zip(("classify", "setup", "tma+tables", "search"), d.T):
                print(f"   {name:12s} cycles: median {np.median(col):8.0f}  p90 {np.percentile(col,90):8.0f}  max {col.max():8.0f}")
            print(f"   active CTA duration ns: median {np.median(dur_ns):.0f} p90 {np.percentile(dur_ns,90):.0f} max {dur_ns.max():.0f}")
            items = t[act][:, 7]
            print(f"   items per active CTA: median {np.median(items):.0f} p90 {np.percentile(items,90):.0f} max {items.max()}")
            # per-SM busy time
            sm = t[:, 6]
            busy = {}
            for smid in np.unique(sm):
                sel = sm == smid
                busy[smid] = (end[sel].max() - start[sel].min()) / 1e3
            b = np.array(list(busy.values()))
            print(f"   per-SM first-start..last-end us: min {b.min():.1f} median {np.median(b):.1f} max {b.max():.1f}; CTAs/SM {len(t)/len(b):.1f}")
            order = np.argsort(-(t[:, 5] - t[:, 0]))[:8]
            for i in order:
                tl = ids[i]
                print(f"   slow CTA tile {tl:4d} (x {tl % 20:2d}, y {tl // 20:2d}) sm {t[i,6]:3d}: start {start[i]/1e3:5.1f} total {(t[i,5]-t[i,0])/1e3:6.1f} us; "
                      f"classify {t[i,1]:6d} setup {t[i,2]-t[i,1]:6d} tma {t[i,3]-t[i,2]:6d} search {t[i,4]-t[i,3]:6d} cycles; items {t[i,7]} "
                      f"seeds {t[i,12]} zeff {t[i,13] & 255} sparse {(t[i,13] >> 8) & 1} cands strip {t[i,8]} global {t[i,9]} "
                      f"warp passes strip {(t[i,13] >> 16) & 0xffffff} global {t[i,13] >> 40} rounds {t[i,14] >> 40} "
                      f"decode cyc/round {(t[i,14] & ((1 << 40) - 1)) / max(1, t[i,14] >> 40):.0f} cand cyc/round {t[i,15] / max(1, t[i,14] >> 40):.0f} "
                      f"bbox {t[i,10] & 0xffff}x{t[i,10] >> 16} strip {t[i,11] & 0xffff}x{t[i,11] >> 16}")
            print(f"   candidates scored from strip {t[:,8].sum()}, from global (fallback) {t[:,9].sum()}; sparse tiles {((t[:,13] >> 8) & 1).sum()}, split tiles {((t[:,13] & 255) > 1).sum()}")
            inact = ~act
            if inact.any():
                di = (t[:, 5] - t[:, 0])[inact] if False else None
