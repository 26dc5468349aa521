"""Generates tests/golden/undistort_cv2.npz with the REAL OpenCV (cv2, 4.13.0 in this image):
what rmd::Depthmap::initUndistortionMap + inputImage compute (src/depthmap.cpp:45-61,95-106):

    cv::initUndistortRectifyMap(K, (k1 k2 r1 r2), I, K, size, CV_16SC2, map1, map2)
    cv::remap(img_8uc1, undistorted_8uc1, map1, map2, INTER_LINEAR)
    undistorted_8uc1.convertTo(img_32fc1, CV_32F, 1.0f/255.0f)

for three cameras:
  "px4"   the reference's own parameters (launch/px4_2.launch), 752x480 -- too large to commit in
          full, so the fixture holds SHA-256 digests of map1 / map2 / remapped image / float image
          and a 64x48 crop of each from the top-left corner (strong barrel distortion: the crop
          includes pixels that sample outside the source);
  "small" the same lens scaled to 188x120, stored in full;
  "tilt"  188x120 with large tangential terms and a negative fy (the data set's convention), in full.
Run where cv2 is importable:   python tests/golden/make_golden_undistort.py
CPU tests pin oracle/rmd_oracle_ingest.c to these vectors, GPU tests pin the product.
"""
import hashlib
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def texture(w, h, seed):
    """8-bit test pattern from integer arithmetic only (identical on every numpy): smooth bands + hash noise."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    hsh = (xx * 73856093) ^ (yy * 19349663) ^ (seed * 83492791)
    hsh = (hsh ^ (hsh >> 13)) * 1274126177 & 0xFFFFFFFF
    noise = (hsh >> 16) % 61                                        # 0..60
    bands = ((xx * 7 + yy * 3) % 256 + (xx * yy // 17) % 128) // 2  # 0..191
    return np.clip(bands + noise, 0, 255).astype(np.uint8)


def run(w, h, fx, fy, cx, cy, k1, k2, r1, r2, seed):
    f32 = np.float32
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], f32)   # cv_K_ is a Mat_<float>, src/depthmap.cpp:35
    D = np.array([[k1, k2, r1, r2]], f32)                      # cv_D_, :50
    m1, m2 = cv2.initUndistortRectifyMap(K, D, np.eye(3), K, (w, h), cv2.CV_16SC2)
    img = texture(w, h, seed)
    und = cv2.remap(img, m1, m2, cv2.INTER_LINEAR)
    flt = cv2.multiply(und, 1.0, scale=float(f32(1.0) / f32(255.0)), dtype=cv2.CV_32F)   # == convertTo(CV_32F, 1/255.f)
    return {"size": np.array([w, h], np.int32), "camera": np.array([fx, fy, cx, cy], f32),
            "dist": np.array([k1, k2, r1, r2], f32), "img": img, "map1": m1, "map2": m2, "undistorted": und,
            "float": flt}


out = {"opencv_version": np.array(cv2.__version__)}
px4 = run(752, 480, 418.715779404372, 418.186411716775, 397.573670639476, 246.235858293295,
          -0.294854287021541, 0.0780596214365872, -0.000520874224877783, 9.42576963868232e-06, 1)
for k in ("size", "camera", "dist"):
    out["px4_" + k] = px4[k]
out["px4_img_seed"] = np.array(1)
for k in ("map1", "map2", "undistorted", "float"):
    out[f"px4_{k}_sha256"] = np.array(sha(px4[k]))
    out[f"px4_{k}_crop"] = px4[k][:48, :64].copy()


def outside(d):
    w, h = d["size"]
    m = d["map1"]
    return int(((m[..., 0] < 0) | (m[..., 1] < 0) | (m[..., 0] >= w - 1) | (m[..., 1] >= h - 1)).sum())


s = 0.25
small = run(188, 120, 418.715779404372 * s, 418.186411716775 * s, 397.573670639476 * s, 246.235858293295 * s,
            -0.294854287021541, 0.0780596214365872, -0.000520874224877783, 9.42576963868232e-06, 2)
tilt = run(188, 120, 120.3, -120.0, 93.5, 59.5, 0.31, -0.05, 0.02, -0.015, 3)
for name, d in (("small", small), ("tilt", tilt)):
    for k, v in d.items():
        out[f"{name}_{k}"] = v
np.savez_compressed(os.path.join(HERE, "undistort_cv2.npz"), **out)
print("wrote undistort_cv2.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if "crop" not in k})
for name, d in (("px4", px4), ("small", small), ("tilt", tilt)):
    out[f"{name}_outside_pixels"] = np.array(outside(d))
    print(name, "pixels with a tap outside the source (border constant 0):", outside(d))
np.savez_compressed(os.path.join(HERE, "undistort_cv2.npz"), **out)
