"""ctypes binding of oracle/_ref/librmd_ref*.so: the reference's own,
unmodified CUDA kernels rebuilt for sm_100a (recipe: oracle/Makefile).
Test infrastructure / bench `--impl reference` only."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF_DIR = os.path.join(_ROOT, "oracle", "_ref")
_libs = {}


def lib_path(patch=5):
    return os.path.join(_REF_DIR, "librmd_ref.so" if patch == 5 else "librmd_ref_p%d.so" % patch)


def available(patch=5) -> bool:
    return os.path.exists(lib_path(patch))


def lib(patch=5):
    if patch not in _libs:
        L = ctypes.CDLL(lib_path(patch))
        vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        L.ref_last_error.restype = ctypes.c_char_p
        L.ref_patch_side.restype = ci
        L.ref_device_count.restype = ci
        L.ref_sync.restype = ci
        L.ref_seeds_create.restype = vp
        L.ref_seeds_create.argtypes = [ci, ci, cf, cf, cf, cf]
        L.ref_seeds_destroy.argtypes = [vp]
        L.ref_seeds_set_reference.argtypes = [vp, vp, vp, cf, cf]
        L.ref_seeds_update.argtypes = [vp, vp, vp]
        L.ref_seeds_download.argtypes = [vp, ci, vp]
        L.ref_seeds_upload.argtypes = [vp, ci, vp]
        L.ref_seeds_converged_count.restype = ctypes.c_longlong
        L.ref_seeds_converged_count.argtypes = [vp]
        L.ref_seeds_dist_from_ref.restype = cf
        L.ref_seeds_dist_from_ref.argtypes = [vp]
        L.ref_denoiser_create.restype = vp
        L.ref_denoiser_create.argtypes = [ci, ci]
        L.ref_denoiser_destroy.argtypes = [vp]
        L.ref_denoiser_run.argtypes = [vp, vp, cf, cf, ci, vp]
        L.ref_reduce_sum_f32.argtypes = [vp, ci, ci, ctypes.POINTER(cf)]
        L.ref_reduce_sum_i32.argtypes = [vp, ci, ci, ctypes.POINTER(ci)]
        L.ref_reduce_count_eq_i32.argtypes = [vp, ci, ci, ci, ctypes.POINTER(ctypes.c_longlong)]
        assert L.ref_patch_side() == patch
        _libs[patch] = L
    return _libs[patch]


def _check(L, rc, what):
    if rc != 0:
        raise RuntimeError(f"reference CUDA: {what}: {L.ref_last_error().decode(errors='replace')} ({rc})")


class RefSeeds:
    """The reference's rmd::SeedMatrix, driven through oracle/ref_driver.cu."""

    def __init__(self, width, height, fx, fy, cx, cy, patch=5):
        self.width, self.height, self.patch = width, height, patch
        self._L = lib(patch)
        self._h = self._L.ref_seeds_create(width, height, fx, fy, cx, cy)
        if not self._h:
            raise RuntimeError("ref_seeds_create failed: " + self._L.ref_last_error().decode())

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.ref_seeds_destroy(h)

    def set_reference(self, img, T_curr_world, min_depth, max_depth):
        img = np.ascontiguousarray(img, np.float32)
        T = np.ascontiguousarray(np.asarray(T_curr_world, np.float32).reshape(12))
        _check(self._L, self._L.ref_seeds_set_reference(self._h, img.ctypes.data, T.ctypes.data,
                                                        min_depth, max_depth), "setReferenceImage")

    def update(self, img, T_curr_world):
        img = np.ascontiguousarray(img, np.float32)
        T = np.ascontiguousarray(np.asarray(T_curr_world, np.float32).reshape(12))
        _check(self._L, self._L.ref_seeds_update(self._h, img.ctypes.data, T.ctypes.data), "update")

    def download(self, field):
        if field == 4:
            out = np.empty((self.height, self.width), np.int32)
        elif field == 7:
            out = np.empty((self.height, self.width, 2), np.float32)
        else:
            out = np.empty((self.height, self.width), np.float32)
        _check(self._L, self._L.ref_seeds_download(self._h, field, out.ctypes.data), "download")
        return out

    def upload(self, field, values):
        a = np.ascontiguousarray(values, np.float32)
        _check(self._L, self._L.ref_seeds_upload(self._h, field, a.ctypes.data), "upload")

    def converged_count(self):
        return int(self._L.ref_seeds_converged_count(self._h))

    def dist_from_ref(self):
        return float(self._L.ref_seeds_dist_from_ref(self._h))

    def sync(self):
        self._L.ref_sync()


class RefDenoiser:
    def __init__(self, width, height, patch=5):
        self.width, self.height = width, height
        self._L = lib(patch)
        self._h = self._L.ref_denoiser_create(width, height)
        if not self._h:
            raise RuntimeError("ref_denoiser_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.ref_denoiser_destroy(h)

    def run(self, seeds: RefSeeds, depth_range, lam, iterations):
        out = np.empty((self.height, self.width), np.float32)
        _check(self._L, self._L.ref_denoiser_run(self._h, seeds._h, depth_range, lam, iterations,
                                                 out.ctypes.data), "denoise")
        return out


def reduce_sum_f32(img):
    L = lib(5)
    img = np.ascontiguousarray(img, np.float32)
    out = ctypes.c_float()
    _check(L, L.ref_reduce_sum_f32(img.ctypes.data, img.shape[1], img.shape[0], ctypes.byref(out)), "sum")
    return float(out.value)


def reduce_count_eq(img, value):
    L = lib(5)
    img = np.ascontiguousarray(img, np.int32)
    out = ctypes.c_longlong()
    _check(L, L.ref_reduce_count_eq_i32(img.ctypes.data, img.shape[1], img.shape[0], value,
                                        ctypes.byref(out)), "countEqual")
    return int(out.value)
