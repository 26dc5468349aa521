"""Where does the end-to-end (host buffer in) frame time go?  GPU box only."""
import os, sys, time, ctypes
os.environ.setdefault('RMD_HOST_PROFILE', '1')
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rpg_open_remode_b200 as rmd
from rpg_open_remode_b200 import synth, _native

W, H, N = 640, 480, 200
seq = synth.SyntheticSequence(W, H, seed=0x5EED0002)
frames = []; poses = []
for k in range(N):
    f = seq.frame(k, want_depth=(k == 0))
    frames.append(f.image); poses.append(f.T_cam_world.reshape(12).copy())
    if k == 0: dmin, dmax = float(f.depth.min()), float(f.depth.max())
u8 = [(f * 255.0 + 0.5).astype(np.uint8) for f in frames]

# raw H2D bandwidth pinned -> device
a = torch.empty(W * H, dtype=torch.float32).pin_memory(); b = torch.empty(W * H, dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): b.copy_(a, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 200
print("pinned H2D 1.2 MB: %.1f us (%.1f GB/s)" % (dt * 1e6, W * H * 4 / dt / 1e9))
src = np.empty(W * H, np.float32); dst = np.empty(W * H, np.float32)
t0 = time.perf_counter()
for _ in range(200): np.copyto(dst, src)
print("numpy memcpy 1.2 MB: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))

def run(label, variant, use_u8=False):
    s = rmd.SeedMatrix(W, H, rmd.PinholeCamera(*seq.camera))
    s.setOption(rmd.OPT_KERNEL_VARIANT, variant)
    src_frames = u8 if use_u8 else frames
    for rep in range(3):
        s.setReferenceImage(src_frames[0], poses[0], dmin, dmax)
        s.sync()
        call = []
        t0 = time.perf_counter()
        for k in range(1, N):
            c0 = time.perf_counter()
            s.update(src_frames[k], poses[k])
            call.append(time.perf_counter() - c0)
        t_enq = time.perf_counter() - t0
        s.sync()
        t_all = time.perf_counter() - t0
    call = np.array(call) * 1e6
    prof = (ctypes.c_double * 8)()
    _native.lib().rmd_debug_host_profile(prof, 1)
    if prof[6] > 0:
        n = prof[6]
        print("    host profile per call (us): wait-slot %.1f  pinned-copy %.1f  h2d-enqueue %.1f  tma-encode %.1f  launch %.1f  total %.1f  (%d calls)" %
              (*(prof[i] / n * 1e6 for i in range(6)), int(n)))
    print("%-34s total %.2f ms (%.0f fps)  enqueue %.2f ms  per-call median %.1f us p90 %.1f max %.1f" %
          (label, t_all * 1e3, (N - 1) / t_all, t_enq * 1e3, np.median(call), np.percentile(call, 90), call.max()))

run("float, staged", rmd.VARIANT_STAGED)
run("float, direct", rmd.VARIANT_DIRECT)
run("u8, staged", rmd.VARIANT_STAGED, use_u8=True)
