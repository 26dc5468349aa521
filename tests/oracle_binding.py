"""ctypes binding of oracle/librmd_oracle.so -- the CHECKER.

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_LIB_PATH = os.path.join(_ORACLE_DIR, "librmd_oracle.so")

UPDATE, CONVERGED, BORDER, DIVERGED, NO_MATCH, NOT_VISIBLE = range(6)
F_MU, F_SIGMA_SQ, F_A, F_B, F_CONVERGENCE, F_SUM_TEMPL, F_CONST_TEMPL_DENOM, F_MATCHES, F_REF_IMG = range(9)

_lib = None


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(_ORACLE_DIR, f) for f in os.listdir(_ORACLE_DIR) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "cpu"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = ctypes.CDLL(_LIB_PATH)
        vp, ci, cf, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
        L.rmd_oracle_seeds_create.restype = vp
        L.rmd_oracle_seeds_create.argtypes = [ci, ci, cf, cf, cf, cf, ci]
        L.rmd_oracle_seeds_destroy.argtypes = [vp]
        L.rmd_oracle_seeds_set_tex_model.argtypes = [vp, ci]
        L.rmd_oracle_set_threads.argtypes = [ci]
        L.rmd_oracle_get_threads.restype = ci
        L.rmd_oracle_seeds_set_reference.argtypes = [vp, vp, vp, cf, cf]
        L.rmd_oracle_seeds_update.argtypes = [vp, vp, vp]
        L.rmd_oracle_stage_check.argtypes = [vp]
        L.rmd_oracle_stage_match.argtypes = [vp, vp, vp]
        L.rmd_oracle_stage_update.argtypes = [vp, vp]
        L.rmd_oracle_seeds_field.restype = vp
        L.rmd_oracle_seeds_field.argtypes = [vp, ci]
        L.rmd_oracle_seeds_converged_count.restype = cs
        L.rmd_oracle_seeds_converged_count.argtypes = [vp]
        L.rmd_oracle_seeds_dist_from_ref.restype = cf
        L.rmd_oracle_seeds_dist_from_ref.argtypes = [vp]
        L.rmd_oracle_seeds_T_curr_ref.argtypes = [vp, vp]
        L.rmd_oracle_denoise.restype = ci
        L.rmd_oracle_denoise.argtypes = [vp, vp, vp, vp, ci, ci, cf, cf, ci, vp]
        L.rmd_oracle_sum_f32_ref_order.restype = cf
        L.rmd_oracle_sum_f32_ref_order.argtypes = [vp, cs, cs, cs]
        L.rmd_oracle_sum_f32_f64.restype = ctypes.c_double
        L.rmd_oracle_sum_f32_f64.argtypes = [vp, cs, cs, cs]
        L.rmd_oracle_sum_i32.restype = ci
        L.rmd_oracle_sum_i32.argtypes = [vp, cs, cs, cs]
        L.rmd_oracle_count_equal_i32.restype = cs
        L.rmd_oracle_count_equal_i32.argtypes = [vp, cs, cs, cs, ci]
        L.rmd_oracle_se3_inv.argtypes = [vp, vp]
        L.rmd_oracle_se3_mul.argtypes = [vp, vp, vp]
        L.rmd_oracle_se3_from_quat.argtypes = [cf] * 7 + [vp]
        L.rmd_oracle_undistort_maps.argtypes = [ci, ci, cf, cf, cf, cf, cf, cf, cf, cf, vp, vp]
        L.rmd_oracle_remap_u8.argtypes = [vp, ci, ci, vp, vp, vp]
        L.rmd_oracle_u8_to_float.argtypes = [vp, cs, vp]
        L.rmd_oracle_point_cloud.argtypes = [vp, vp, vp, ci, ci, cf, cf, cf, cf, vp, vp]
        L.rmd_oracle_point_cloud.restype = cs
        _lib = L
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def se3_inv(T):
    T = _f32(T).reshape(12)
    out = np.empty(12, np.float32)
    lib().rmd_oracle_se3_inv(T.ctypes.data, out.ctypes.data)
    return out.reshape(3, 4)


def se3_mul(A, B):
    A, B = _f32(A).reshape(12), _f32(B).reshape(12)
    out = np.empty(12, np.float32)
    lib().rmd_oracle_se3_mul(A.ctypes.data, B.ctypes.data, out.ctypes.data)
    return out.reshape(3, 4)


def se3_from_quat(qw, qx, qy, qz, tx, ty, tz):
    out = np.empty(12, np.float32)
    lib().rmd_oracle_se3_from_quat(qw, qx, qy, qz, tx, ty, tz, out.ctypes.data)
    return out.reshape(3, 4)


class OracleSeeds:
    """CPU restatement of rmd::SeedMatrix (same call sequence)."""

    _DTYPES = {F_CONVERGENCE: np.int32}

    def __init__(self, width, height, fx, fy, cx, cy, patch=5, tex_frac_bits=8):
        self.width, self.height, self.patch = width, height, patch
        self._L = lib()
        self._h = self._L.rmd_oracle_seeds_create(width, height, fx, fy, cx, cy, patch)
        if not self._h:
            raise ValueError("rmd_oracle_seeds_create failed")
        self._L.rmd_oracle_seeds_set_tex_model(self._h, tex_frac_bits)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.rmd_oracle_seeds_destroy(h)

    def set_reference(self, img, T_curr_world, min_depth, max_depth):
        img, T = _f32(img), _f32(T_curr_world).reshape(12)
        assert img.shape == (self.height, self.width)
        self._L.rmd_oracle_seeds_set_reference(self._h, img.ctypes.data, T.ctypes.data, min_depth, max_depth)

    def update(self, img, T_curr_world):
        img, T = _f32(img), _f32(T_curr_world).reshape(12)
        assert img.shape == (self.height, self.width)
        self._L.rmd_oracle_seeds_update(self._h, img.ctypes.data, T.ctypes.data)

    def stage_check(self):
        self._L.rmd_oracle_stage_check(self._h)

    def stage_match(self, img, T_curr_ref):
        img, T = _f32(img), _f32(T_curr_ref).reshape(12)
        self._L.rmd_oracle_stage_match(self._h, img.ctypes.data, T.ctypes.data)

    def stage_update(self, T_ref_curr):
        T = _f32(T_ref_curr).reshape(12)
        self._L.rmd_oracle_stage_update(self._h, T.ctypes.data)

    def field(self, f):
        """Live numpy view of the oracle's state (writes go through)."""
        n = self.width * self.height
        ptr = self._L.rmd_oracle_seeds_field(self._h, f)
        if f == F_MATCHES:
            buf = (ctypes.c_float * (2 * n)).from_address(ptr)
            return np.frombuffer(buf, dtype=np.float32).reshape(self.height, self.width, 2)
        if f == F_CONVERGENCE:
            buf = (ctypes.c_int32 * n).from_address(ptr)
            return np.frombuffer(buf, dtype=np.int32).reshape(self.height, self.width)
        buf = (ctypes.c_float * n).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float32).reshape(self.height, self.width)

    mu = property(lambda s: s.field(F_MU))
    sigma_sq = property(lambda s: s.field(F_SIGMA_SQ))
    a = property(lambda s: s.field(F_A))
    b = property(lambda s: s.field(F_B))
    convergence = property(lambda s: s.field(F_CONVERGENCE))
    sum_templ = property(lambda s: s.field(F_SUM_TEMPL))
    const_templ_denom = property(lambda s: s.field(F_CONST_TEMPL_DENOM))
    matches = property(lambda s: s.field(F_MATCHES))

    def converged_count(self):
        return int(self._L.rmd_oracle_seeds_converged_count(self._h))

    def dist_from_ref(self):
        return float(self._L.rmd_oracle_seeds_dist_from_ref(self._h))

    def T_curr_ref(self):
        out = np.empty(12, np.float32)
        self._L.rmd_oracle_seeds_T_curr_ref(self._h, out.ctypes.data)
        return out.reshape(3, 4)


def denoise(mu, sigma_sq, a, b, depth_range, lam, iterations):
    mu, sigma_sq, a, b = map(_f32, (mu, sigma_sq, a, b))
    h, w = mu.shape
    out = np.empty((h, w), np.float32)
    ok = lib().rmd_oracle_denoise(mu.ctypes.data, sigma_sq.ctypes.data, a.ctypes.data, b.ctypes.data,
                                  w, h, depth_range, lam, iterations, out.ctypes.data)
    assert ok
    return out


def sum_f32_ref_order(img):
    img = _f32(img)
    h, w = img.shape
    return float(lib().rmd_oracle_sum_f32_ref_order(img.ctypes.data, w, w, h))


def sum_f32_f64(img):
    img = _f32(img)
    h, w = img.shape
    return float(lib().rmd_oracle_sum_f32_f64(img.ctypes.data, w, w, h))


def sum_i32(img):
    img = np.ascontiguousarray(img, np.int32)
    h, w = img.shape
    return int(lib().rmd_oracle_sum_i32(img.ctypes.data, w, w, h))


def count_equal_i32(img, value):
    img = np.ascontiguousarray(img, np.int32)
    h, w = img.shape
    return int(lib().rmd_oracle_count_equal_i32(img.ctypes.data, w, w, h, value))


def set_threads(n):
    lib().rmd_oracle_set_threads(n)


def get_threads():
    return lib().rmd_oracle_get_threads()


# ---- frame ingest (oracle/rmd_oracle_ingest.c): src/depthmap.cpp:45-61,95-106
def undistort_maps(width, height, fx, fy, cx, cy, k1, k2, p1, p2):
    """(map1 int16 [h, w, 2], map2 uint16 [h, w]) of cv::initUndistortRectifyMap(..., CV_16SC2)."""
    m1 = np.empty((height, width, 2), np.int16)
    m2 = np.empty((height, width), np.uint16)
    lib().rmd_oracle_undistort_maps(width, height, fx, fy, cx, cy, k1, k2, p1, p2, m1.ctypes.data, m2.ctypes.data)
    return m1, m2


def remap_u8(img, map1, map2):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    m1, m2 = np.ascontiguousarray(map1, np.int16), np.ascontiguousarray(map2, np.uint16)
    out = np.empty_like(img)
    lib().rmd_oracle_remap_u8(img.ctypes.data, w, h, m1.ctypes.data, m2.ctypes.data, out.ctypes.data)
    return out


def u8_to_float(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(img.shape, np.float32)
    lib().rmd_oracle_u8_to_float(img.ctypes.data, img.size, out.ctypes.data)
    return out


# ---- point cloud (oracle/rmd_oracle_pointcloud.c): src/publisher.cpp:54-86
def point_cloud(depth, conv, ref_u8, fx, fy, cx, cy, T_world_ref):
    """float32 [n, 4] = (x, y, z, intensity) of the CONVERGED pixels, row-major order."""
    depth = np.ascontiguousarray(depth, np.float32)
    conv = np.ascontiguousarray(conv, np.int32)
    ref_u8 = np.ascontiguousarray(ref_u8, np.uint8)
    h, w = depth.shape
    T = np.ascontiguousarray(np.asarray(T_world_ref, np.float32).reshape(-1)[:12])
    n = lib().rmd_oracle_point_cloud(depth.ctypes.data, conv.ctypes.data, ref_u8.ctypes.data, w, h, fx, fy, cx, cy,
                                     T.ctypes.data, None)
    out = np.empty((n, 4), np.float32)
    lib().rmd_oracle_point_cloud(depth.ctypes.data, conv.ctypes.data, ref_u8.ctypes.data, w, h, fx, fy, cx, cy,
                                 T.ctypes.data, out.ctypes.data)
    return out
