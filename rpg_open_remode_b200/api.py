"""Host-side mirror of the reference's C++ interface for the hot path.

Same class and method names, argument meaning and error behaviour as

* ``rmd::SeedMatrix``        include/rmd/seed_matrix.cuh:45-109
* ``rmd::DepthmapDenoiser``  include/rmd/depthmap_denoiser.cuh:27-54
* ``rmd::ImageReducer<T>``   include/rmd/reduction.cuh:27-62
* ``rmd::DeviceImage<T>``    include/rmd/device_image.cuh:34-180
* ``rmd::Depthmap``          include/rmd/depthmap.h:34-129 (OpenCV-free)
* ``rmd::SE3<float>``, ``rmd::PinholeCamera``

so the parity tests read like the reference's gtests.  Every call goes
through the C-ABI (``include/rmd_b200.h``) into the sm_100a kernels; nothing
is computed in Python and there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from . import _native
from ._native import RmdError, check

# rmd::ConvergenceStates, include/rmd/seed_matrix.cuh:31-43
class ConvergenceStates:
    UPDATE = 0
    CONVERGED = 1
    BORDER = 2
    DIVERGED = 3
    NO_MATCH = 4
    NOT_VISIBLE = 5


FIELD_MU, FIELD_SIGMA_SQ, FIELD_A, FIELD_B, FIELD_CONVERGENCE = 0, 1, 2, 3, 4
FIELD_SUM_TEMPL, FIELD_CONST_TEMPL_DENOM, FIELD_EPIPOLAR_MATCHES, FIELD_REF_IMG = 5, 6, 7, 8
OPT_RECORD_MATCHES, OPT_KERNEL_VARIANT, OPT_TEX_FRAC_BITS, OPT_DEBUG_TIMELINE, OPT_PINNED_INPUT = 0, 1, 2, 3, 4
OPT_CHAIN_FRAMES, OPT_SEED_MODE_PCT = 5, 6
# tuning knobs of the staged kernel (include/rmd_b200.h RMD_OPT_TUNE_*; results never depend on them)
(OPT_TUNE_SPLIT_MAX, OPT_TUNE_SPLIT_MIN_ITEMS, OPT_TUNE_SPLIT_ITEMS_PER_CTA, OPT_TUNE_SPARSE_MAX_SEEDS,
 OPT_TUNE_HEAVY_MIN_ITEMS, OPT_TUNE_SPLIT_AVG_PCT, OPT_TUNE_PDL) = 10, 11, 12, 13, 14, 15, 16
OPT_TUNE_WARP_TILE_SEEDS = 17
OPT_TUNE_GRID_CTAS = 18
OPT_TUNE_CTAS_PER_SM = 19
OPT_TUNE_WARP_TILE_CANDS = 20
FIELD_DEBUG_TIMELINE = 100
VARIANT_STAGED, VARIANT_DIRECT = 0, 1

_f32 = np.float32


class PinholeCamera:
    """include/rmd/pinhole_camera.cuh:27-63"""

    def __init__(self, fx=0.0, fy=0.0, cx=0.0, cy=0.0):
        self.fx, self.fy, self.cx, self.cy = (float(_f32(v)) for v in (fx, fy, cx, cy))

    def cam2world(self, uv):
        u, v = _f32(uv[0]), _f32(uv[1])
        return np.array([(u - _f32(self.cx)) / _f32(self.fx), (v - _f32(self.cy)) / _f32(self.fy), 1.0], _f32)

    def world2cam(self, xyz):
        x, y, z = (_f32(t) for t in xyz)
        return np.array([_f32(self.fx) * x / z + _f32(self.cx), _f32(self.fy) * y / z + _f32(self.cy)], _f32)

    def getOnePixAngle(self):
        return float(_f32(math.atan2(1.0, 2.0 * self.fx)) * _f32(2.0))


class SE3:
    """include/rmd/se3.cuh:27-168: 3x4 row-major [R|t], float32 arithmetic."""

    def __init__(self, *args):
        if len(args) == 0:
            self.data = np.zeros(12, _f32)
            self.data[[0, 5, 10]] = 1.0
        elif len(args) == 1:
            self.data = np.array(args[0], dtype=_f32).reshape(12).copy()
        elif len(args) == 2:  # (r row-major 3x3, t)
            r, t = np.asarray(args[0], _f32).reshape(3, 3), np.asarray(args[1], _f32).reshape(3)
            self.data = np.concatenate([r, t[:, None]], axis=1).reshape(12).astype(_f32)
        elif len(args) == 7:  # (qw, qx, qy, qz, tx, ty, tz), se3.cuh:37-66
            qw, qx, qy, qz, tx, ty, tz = (_f32(a) for a in args)
            two = _f32(2)
            x, y, z = two * qx, two * qy, two * qz
            wx, wy, wz = x * qw, y * qw, z * qw
            xx, xy, xz = x * qx, y * qx, z * qx
            yy, yz, zz = y * qy, z * qy, z * qz
            one = _f32(1)
            self.data = np.array([one - (yy + zz), xy - wz, xz + wy, tx,
                                  xy + wz, one - (xx + zz), yz - wx, ty,
                                  xz - wy, yz + wx, one - (xx + yy), tz], _f32)
        else:
            raise TypeError("SE3(): expected (), (12 floats), (r, t) or (qw,qx,qy,qz,tx,ty,tz)")

    def __call__(self, r, c):
        return float(self.data[4 * r + c])

    def inv(self):
        d, o = self.data, np.empty(12, _f32)
        o[0], o[1], o[2] = d[0], d[4], d[8]
        o[4], o[5], o[6] = d[1], d[5], d[9]
        o[8], o[9], o[10] = d[2], d[6], d[10]
        o[3] = -d[0] * d[3] - d[4] * d[7] - d[8] * d[11]
        o[7] = -d[1] * d[3] - d[5] * d[7] - d[9] * d[11]
        o[11] = -d[2] * d[3] - d[6] * d[7] - d[10] * d[11]
        return SE3(o)

    def __mul__(self, other):
        if isinstance(other, SE3):
            l, r, o = self.data, other.data, np.empty(12, _f32)
            for row in range(3):
                a = l[4 * row:4 * row + 4]
                for col in range(3):
                    o[4 * row + col] = a[0] * r[col] + a[1] * r[4 + col] + a[2] * r[8 + col]
                o[4 * row + 3] = a[3] + a[0] * r[3] + a[1] * r[7] + a[2] * r[11]
            return SE3(o)
        return self.translate(self.rotate(other))

    def rotate(self, p):
        d, p = self.data, np.asarray(p, _f32)
        return np.array([d[0] * p[0] + d[1] * p[1] + d[2] * p[2],
                         d[4] * p[0] + d[5] * p[1] + d[6] * p[2],
                         d[8] * p[0] + d[9] * p[1] + d[10] * p[2]], _f32)

    def translate(self, p):
        p = np.asarray(p, _f32)
        return np.array([p[0] + self.data[3], p[1] + self.data[7], p[2] + self.data[11]], _f32)

    def getTranslation(self):
        return self.data[[3, 7, 11]].copy()

    def __repr__(self):
        return "SE3(\n%s)" % self.data.reshape(3, 4)


def _pose12(T) -> np.ndarray:
    if isinstance(T, SE3):
        return np.ascontiguousarray(T.data, dtype=_f32)
    a = np.ascontiguousarray(np.asarray(T, dtype=_f32).reshape(-1))
    if a.size != 12:
        raise ValueError("pose must be SE3 or 12 floats (3x4 row-major)")
    return a


class DeviceImage:
    """include/rmd/device_image.cuh:34-180.  ``dtype`` in {float32, int32,
    'float2'}.  Public fields as in the reference: width, height, pitch,
    stride (elements), data (device address)."""

    _ELEM = {"float32": (4, np.float32, 1), "int32": (4, np.int32, 1), "float2": (8, np.float32, 2)}

    def __init__(self, width, height, dtype="float32", _view=None):
        key = "float2" if dtype == "float2" else np.dtype(dtype).name
        if key not in self._ELEM:
            raise TypeError(f"DeviceImage: unsupported element type {dtype}")
        self._elem_size, self._np, self._comps = self._ELEM[key]
        self.dtype = key
        self.width, self.height = int(width), int(height)
        self._L = _native.lib()
        if _view is not None:
            self.data, self.pitch = int(_view[0]), int(_view[1])
            self._owned = False
        else:
            ptr, pitch = ctypes.c_void_p(), ctypes.c_size_t()
            check(self._L.rmd_image_alloc(self.width, self.height, self._elem_size,
                                          ctypes.byref(ptr), ctypes.byref(pitch)),
                  "Image: unable to allocate pitched memory.")
            self.data, self.pitch = ptr.value, pitch.value
            self._owned = True
        self.stride = self.pitch // self._elem_size

    def __del__(self):
        if getattr(self, "_owned", False) and getattr(self, "data", None):
            self._L.rmd_image_free(self.data)
            self.data = None

    def _host_shape(self):
        return (self.height, self.width) if self._comps == 1 else (self.height, self.width, self._comps)

    def setDevData(self, aligned_data_row_major):
        a = np.ascontiguousarray(aligned_data_row_major, dtype=self._np)
        if a.shape != self._host_shape():
            raise ValueError(f"setDevData: expected shape {self._host_shape()}, got {a.shape}")
        check(self._L.rmd_image_upload(self.data, self.pitch, a.ctypes.data, self.width, self.height,
                                       self._elem_size), "Image: unable to copy data from host to device.")

    def getDevData(self):
        out = np.empty(self._host_shape(), dtype=self._np)
        check(self._L.rmd_image_download(self.data, self.pitch, out.ctypes.data, self.width, self.height,
                                         self._elem_size), "Image: unable to copy data from device to host.")
        return out

    def zero(self):
        check(self._L.rmd_image_zero(self.data, self.pitch, self.width, self.height, self._elem_size),
              "Image: unable to zero.")

    def assign(self, other: "DeviceImage"):
        """operator= (device to device copy), device_image.cuh:150-171"""
        assert (self.width, self.height, self.dtype) == (other.width, other.height, other.dtype)
        check(self._L.rmd_image_copy(self.data, self.pitch, other.data, other.pitch, self.width,
                                     self.height, self._elem_size),
              "Image, operator '=': unable to copy data from another image.")
        return self


class SeedMatrix:
    """rmd::SeedMatrix -- include/rmd/seed_matrix.cuh:45-109, src/seed_matrix.cu."""

    def __init__(self, width, height, cam: PinholeCamera, patch_side=5, device=-1):
        self.width_, self.height_, self.patch_side = int(width), int(height), int(patch_side)
        self._L = _native.lib()
        h = ctypes.c_void_p()
        check(self._L.rmd_seeds_create(self.width_, self.height_, cam.fx, cam.fy, cam.cx, cam.cy,
                                       self.patch_side, int(device), ctypes.byref(h)), "SeedMatrix")
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.rmd_seeds_destroy(h)

    @property
    def handle(self):
        return self._h

    # ---- reference API
    def setReferenceImage(self, host_ref_img_align_row_maj, T_curr_world, min_depth, max_depth) -> bool:
        img = self._frame(host_ref_img_align_row_maj)
        T = _pose12(T_curr_world)
        fn = self._L.rmd_seeds_set_reference_u8 if img.dtype == np.uint8 else self._L.rmd_seeds_set_reference
        check(fn(self._h, img.ctypes.data, T.ctypes.data, float(min_depth), float(max_depth)),
              "SeedMatrix::setReferenceImage")
        return True

    def update(self, host_curr_img_align_row_maj, T_curr_world) -> bool:
        img = self._frame(host_curr_img_align_row_maj)
        T = _pose12(T_curr_world)
        fn = self._L.rmd_seeds_update_u8 if img.dtype == np.uint8 else self._L.rmd_seeds_update
        check(fn(self._h, img.ctypes.data, T.ctypes.data), "SeedMatrix::update")
        return True

    def downloadDepthmap(self):
        return self._download(FIELD_MU)

    def downloadConvergence(self):
        return self._download(FIELD_CONVERGENCE)

    def getMu(self):
        return self._device_image(FIELD_MU)

    def getSigmaSq(self):
        return self._device_image(FIELD_SIGMA_SQ)

    def getA(self):
        return self._device_image(FIELD_A)

    def getB(self):
        return self._device_image(FIELD_B)

    def getConvergence(self):
        return self._device_image(FIELD_CONVERGENCE)

    def getConvergedCount(self) -> int:
        n = ctypes.c_size_t()
        check(self._L.rmd_seeds_converged_count(self._h, ctypes.byref(n)), "SeedMatrix::getConvergedCount")
        return int(n.value)

    def getDistFromRef(self) -> float:
        d = ctypes.c_float()
        check(self._L.rmd_seeds_dist_from_ref(self._h, ctypes.byref(d)), "SeedMatrix::getDistFromRef")
        return float(d.value)

    # RMD_BUILD_TESTS accessors, seed_matrix.cuh:76-83
    def downloadSigmaSq(self):
        return self._download(FIELD_SIGMA_SQ)

    def downloadA(self):
        return self._download(FIELD_A)

    def downloadB(self):
        return self._download(FIELD_B)

    def downloadSumTempl(self):
        return self._download(FIELD_SUM_TEMPL)

    def downloadConstTemplDenom(self):
        return self._download(FIELD_CONST_TEMPL_DENOM)

    def downloadEpipolarMatches(self):
        return self._download(FIELD_EPIPOLAR_MATCHES)

    # ---- additions of this implementation
    def updateDevice(self, dev_ptr: int, pitch_bytes: int, T_curr_world) -> bool:
        """Frame already resident in device memory (16-byte aligned)."""
        T = _pose12(T_curr_world)
        check(self._L.rmd_seeds_update_device(self._h, dev_ptr, pitch_bytes, T.ctypes.data),
              "SeedMatrix::updateDevice")
        return True

    def updateDeviceBatch(self, dev_ptr: int, frame_stride_bytes: int, pitch_bytes: int, poses) -> bool:
        """n consecutive updates from device-resident frames; poses: (n, 12) float32."""
        T = np.ascontiguousarray(np.asarray(poses, dtype=_f32).reshape(-1, 12))
        check(self._L.rmd_seeds_update_device_batch(self._h, dev_ptr, frame_stride_bytes, pitch_bytes,
                                                    T.shape[0], T.ctypes.data),
              "SeedMatrix::updateDeviceBatch")
        return True

    def setReferenceImageDevice(self, dev_ptr: int, pitch_bytes: int, T_curr_world, min_depth, max_depth):
        T = _pose12(T_curr_world)
        check(self._L.rmd_seeds_set_reference_device(self._h, dev_ptr, pitch_bytes, T.ctypes.data,
                                                     float(min_depth), float(max_depth)),
              "SeedMatrix::setReferenceImageDevice")
        return True

    # ---- frame ingest with lens undistortion (rmd::Depthmap::initUndistortionMap / inputImage,
    # src/depthmap.cpp:45-61,95-106); applies to uint8 frames given to setReferenceImage / update
    def initUndistortionMap(self, k1, k2, r1, r2) -> None:
        check(self._L.rmd_seeds_init_undistortion_map(self._h, float(k1), float(k2), float(r1), float(r2)),
              "SeedMatrix::initUndistortionMap")

    def clearUndistortionMap(self) -> None:
        check(self._L.rmd_seeds_clear_undistortion_map(self._h), "SeedMatrix::clearUndistortionMap")

    def getUndistortionMap(self):
        """(map1, map2) as cv::initUndistortRectifyMap(..., CV_16SC2) lays them out."""
        m1 = np.empty((self.height_, self.width_, 2), np.int16)
        m2 = np.empty((self.height_, self.width_), np.uint16)
        check(self._L.rmd_seeds_get_undistortion_map(self._h, m1.ctypes.data, m2.ctypes.data),
              "SeedMatrix::getUndistortionMap")
        return m1, m2

    def undistort(self, img_8uc1):
        """The remapped 8-bit frame (img_undistorted_8uc1_ of rmd::Depthmap)."""
        a = np.ascontiguousarray(img_8uc1, dtype=np.uint8)
        if a.shape != (self.height_, self.width_):
            raise ValueError("undistort: wrong shape")
        out = np.empty_like(a)
        check(self._L.rmd_seeds_undistort_u8(self._h, a.ctypes.data, out.ctypes.data), "SeedMatrix::undistort")
        return out

    @staticmethod
    def updateMany(seeds: "list[SeedMatrix]", host_curr_img_align_row_maj, T_curr_world) -> bool:
        """Several live keyframes against one incoming frame (rmd_seeds_update_many): one upload, one fused kernel
        per keyframe on its own stream.  Same result as update() on each; an 8-bit frame goes through seeds[0]'s
        ingest (undistortion map)."""
        if not seeds:
            raise ValueError("updateMany: no keyframes")
        first = seeds[0]
        img = first._frame(host_curr_img_align_row_maj)
        T = _pose12(T_curr_world)
        arr = (ctypes.c_void_p * len(seeds))(*[s._h.value for s in seeds])
        fn = first._L.rmd_seeds_update_many_u8 if img.dtype == np.uint8 else first._L.rmd_seeds_update_many
        check(fn(arr, len(seeds), img.ctypes.data, T.ctypes.data), "SeedMatrix::updateMany")
        return True

    def pointCloud(self, depth: "DeviceImage | None" = None, capacity: "int | None" = None):
        """rmd::Publisher::publishPointCloud (src/publisher.cpp:54-86): float32 [n, 4] = (x, y, z, intensity) of the
        CONVERGED pixels in row-major order, from the seeds' own depth or from a device depth image (e.g. denoised).
        Returns (points, count); count > len(points) when `capacity` was too small."""
        cap = self.width_ * self.height_ if capacity is None else int(capacity)
        out = np.empty((cap, 4), np.float32)
        n = ctypes.c_size_t()
        ptr, pitch = (depth.data, depth.pitch) if depth is not None else (None, 0)
        check(self._L.rmd_seeds_point_cloud(self._h, ptr, pitch, out.ctypes.data, cap, ctypes.byref(n)),
              "SeedMatrix::pointCloud")
        return out[:min(cap, n.value)], int(n.value)

    def uploadState(self, field: int, values) -> None:
        dt = np.int32 if field == FIELD_CONVERGENCE else np.float32
        a = np.ascontiguousarray(values, dtype=dt)
        if a.shape != (self.height_, self.width_):
            raise ValueError("uploadState: wrong shape")
        check(self._L.rmd_seeds_upload_state(self._h, field, a.ctypes.data), "SeedMatrix::uploadState")

    def copyFieldToDevice(self, field: int, dev_ptr: int, pitch_bytes: int) -> None:
        check(self._L.rmd_seeds_copy_field_to_device(self._h, field, dev_ptr, pitch_bytes),
              "SeedMatrix::copyFieldToDevice")

    def setOption(self, option: int, value: int) -> None:
        check(self._L.rmd_seeds_set_option(self._h, option, value), "SeedMatrix::setOption")

    def setStream(self, cuda_stream: int) -> None:
        check(self._L.rmd_seeds_set_stream(self._h, cuda_stream), "SeedMatrix::setStream")

    def sync(self) -> None:
        check(self._L.rmd_seeds_sync(self._h), "SeedMatrix::sync")

    def launchCount(self):
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        check(self._L.rmd_seeds_launch_count(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def enableKernelTiming(self, on=True):
        check(self._L.rmd_seeds_enable_kernel_timing(self._h, 1 if on else 0))

    def lastKernelMs(self) -> float:
        ms = ctypes.c_float()
        check(self._L.rmd_seeds_last_kernel_ms(self._h, ctypes.byref(ms)), "SeedMatrix::lastKernelMs")
        return float(ms.value)

    # ---- helpers
    def _frame(self, img):
        a = np.asarray(img)
        if a.dtype != np.uint8:
            a = np.ascontiguousarray(a, dtype=np.float32)
        else:
            a = np.ascontiguousarray(a)
        if a.shape != (self.height_, self.width_):
            raise ValueError(f"frame must be ({self.height_}, {self.width_}), got {a.shape}")
        return a

    def downloadTimeline(self):
        """Debug (OPT_DEBUG_TIMELINE): int64[n_tiles, 16] of the last staged launch."""
        n = ((self.width_ + 31) // 32) * ((self.height_ + 7) // 8)
        out = np.empty((n, 16), np.int64)
        check(self._L.rmd_seeds_download(self._h, FIELD_DEBUG_TIMELINE, out.ctypes.data), "SeedMatrix::downloadTimeline")
        return out

    def _download(self, field):
        if field == FIELD_CONVERGENCE:
            out = np.empty((self.height_, self.width_), np.int32)
        elif field == FIELD_EPIPOLAR_MATCHES:
            out = np.empty((self.height_, self.width_, 2), np.float32)
        else:
            out = np.empty((self.height_, self.width_), np.float32)
        check(self._L.rmd_seeds_download(self._h, field, out.ctypes.data), "SeedMatrix::download")
        return out

    def _device_image(self, field):
        ptr, pitch = ctypes.c_void_p(), ctypes.c_size_t()
        check(self._L.rmd_seeds_device_ptr(self._h, field, ctypes.byref(ptr), ctypes.byref(pitch)),
              "SeedMatrix::get*")
        dt = "int32" if field == FIELD_CONVERGENCE else ("float2" if field == FIELD_EPIPOLAR_MATCHES else "float32")
        return DeviceImage(self.width_, self.height_, dt, _view=(ptr.value, pitch.value))


class DepthmapDenoiser:
    """rmd::DepthmapDenoiser -- include/rmd/depthmap_denoiser.cuh:27-54."""

    def __init__(self, width, height, device=-1):
        self.width, self.height = int(width), int(height)
        self._L = _native.lib()
        h = ctypes.c_void_p()
        check(self._L.rmd_denoiser_create(self.width, self.height, int(device), ctypes.byref(h)),
              "DepthmapDenoiser")
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.rmd_denoiser_destroy(h)

    def setLargeSigmaSq(self, depth_range):
        check(self._L.rmd_denoiser_set_large_sigma_sq(self._h, float(depth_range)))

    def denoise(self, mu: DeviceImage, sigma_sq: DeviceImage, a: DeviceImage, b: DeviceImage,
                lam: float, iterations: int):
        """Returns host_denoised (the reference fills a caller buffer)."""
        out = np.empty((self.height, self.width), np.float32)
        check(self._L.rmd_denoiser_run(self._h, mu.data, mu.pitch, sigma_sq.data, sigma_sq.pitch,
                                       a.data, a.pitch, b.data, b.pitch, out.ctypes.data,
                                       float(lam), int(iterations)), "DepthmapDenoiser::denoise")
        return out

    def denoiseSeeds(self, seeds: SeedMatrix, lam: float, iterations: int):
        out = np.empty((self.height, self.width), np.float32)
        check(self._L.rmd_denoiser_run_seeds(self._h, seeds.handle, out.ctypes.data, float(lam),
                                             int(iterations)), "DepthmapDenoiser::denoiseSeeds")
        return out

    def denoiseSeedsToDevice(self, seeds: SeedMatrix, dev_ptr: int, pitch_bytes: int, lam: float,
                             iterations: int):
        check(self._L.rmd_denoiser_run_seeds_to_device(self._h, seeds.handle, dev_ptr, pitch_bytes,
                                                       float(lam), int(iterations)),
              "DepthmapDenoiser::denoiseSeedsToDevice")

    def setStream(self, cuda_stream: int):
        check(self._L.rmd_denoiser_set_stream(self._h, cuda_stream))

    def sync(self):
        check(self._L.rmd_denoiser_sync(self._h))

    def launchCount(self) -> int:
        n = ctypes.c_uint64()
        check(self._L.rmd_denoiser_launch_count(self._h, ctypes.byref(n)))
        return int(n.value)


class ImageReducer:
    """rmd::ImageReducer<T> -- include/rmd/reduction.cuh:27-62.  The launch
    shape arguments of the reference are accepted and ignored (the kernel
    sizes its own grid)."""

    def __init__(self, dtype="float32", num_threads_per_block=None, num_blocks_per_grid=None):
        self.dtype = np.dtype(dtype).name
        if self.dtype not in ("float32", "int32"):
            raise TypeError("ImageReducer<T>: T must be float32 or int32 (src/reduction.cu:186-187)")
        self._L = _native.lib()

    def sum(self, img: DeviceImage):
        if img.dtype != self.dtype:
            raise TypeError("ImageReducer::sum: element type mismatch")
        if self.dtype == "float32":
            out = ctypes.c_float()
            check(self._L.rmd_reduce_sum_f32(img.data, img.stride, img.width, img.height, ctypes.byref(out)),
                  "sum")
            return float(out.value)
        out = ctypes.c_int32()
        check(self._L.rmd_reduce_sum_i32(img.data, img.stride, img.width, img.height, ctypes.byref(out)), "sum")
        return int(out.value)

    def countEqual(self, img: DeviceImage, value: int) -> int:
        if img.dtype != "int32":
            raise TypeError("countEqual is only instantiated for int (src/reduction.cu:134)")
        out = ctypes.c_size_t()
        check(self._L.rmd_reduce_count_eq_i32(img.data, img.stride, img.width, img.height, int(value),
                                              ctypes.byref(out)), "countEqual")
        return int(out.value)

    def minMax(self, img: DeviceImage):
        lo, hi = ctypes.c_float(), ctypes.c_float()
        check(self._L.rmd_reduce_min_max_f32(img.data, img.stride, img.width, img.height,
                                             ctypes.byref(lo), ctypes.byref(hi)), "minMax")
        return float(lo.value), float(hi.value)


class Depthmap:
    """rmd::Depthmap -- include/rmd/depthmap.h:34-129, src/depthmap.cpp, without
    OpenCV: frames are numpy uint8 (h, w) arrays; the 8U -> 32F * (1/255)
    conversion of inputImage (src/depthmap.cpp:105) runs on the GPU."""

    def __init__(self, width, height, fx, cx, fy, cy, patch_side=5, device=-1):
        self.width_, self.height_ = int(width), int(height)
        self.seeds_ = SeedMatrix(width, height, PinholeCamera(fx, fy, cx, cy), patch_side, device)
        self.denoiser_ = DepthmapDenoiser(width, height, device)
        self.output_depth_32fc1_ = np.zeros((height, width), np.float32)
        self.output_convergence_int_ = np.zeros((height, width), np.int32)
        self.ref_img_undistorted_8uc1_ = np.zeros((height, width), np.uint8)
        self.T_world_ref_ = SE3()
        self.is_distorted_ = False

    def initUndistortionMap(self, k1, k2, r1, r2):
        """src/depthmap.cpp:45-61; the maps and the per-frame remap live on the GPU."""
        self.seeds_.initUndistortionMap(k1, k2, r1, r2)
        self.is_distorted_ = True

    def setReferenceImage(self, img_curr, T_curr_world, min_depth, max_depth) -> bool:
        self.denoiser_.setLargeSigmaSq(max_depth - min_depth)          # src/depthmap.cpp:69
        img = self._input_image(img_curr)
        ret = self.seeds_.setReferenceImage(img, T_curr_world, min_depth, max_depth)
        # img_undistorted_8uc1_.copyTo(ref_img_undistorted_8uc1_), src/depthmap.cpp:78
        self.ref_img_undistorted_8uc1_ = self.seeds_.undistort(img) if self.is_distorted_ else np.array(img, copy=True)
        self.T_world_ref_ = (T_curr_world if isinstance(T_curr_world, SE3) else SE3(T_curr_world)).inv()
        return ret

    def update(self, img_curr, T_curr_world) -> None:
        self.seeds_.update(self._input_image(img_curr), T_curr_world)

    def downloadDepthmap(self) -> None:
        self.output_depth_32fc1_ = self.seeds_.downloadDepthmap()

    def downloadDenoisedDepthmap(self, lam, iterations) -> None:
        self.output_depth_32fc1_ = self.denoiser_.denoiseSeeds(self.seeds_, lam, iterations)

    def getDepthmap(self):
        return self.output_depth_32fc1_

    def downloadConvergenceMap(self) -> None:
        self.output_convergence_int_ = self.seeds_.downloadConvergence()

    def getConvergenceMap(self):
        return self.output_convergence_int_

    def getReferenceImage(self):
        return self.ref_img_undistorted_8uc1_

    def getConvergedCount(self) -> int:
        return self.seeds_.getConvergedCount()

    def getConvergedPercentage(self) -> float:
        return float(self.getConvergedCount()) / float(self.width_ * self.height_) * 100.0

    def getDistFromRef(self) -> float:
        return self.seeds_.getDistFromRef()

    def getWidth(self):
        return self.width_

    def getHeight(self):
        return self.height_

    def getT_world_ref(self):
        return self.T_world_ref_

    def _input_image(self, img_8uc1):
        a = np.asarray(img_8uc1)
        if a.dtype != np.uint8:
            raise TypeError("Depthmap expects 8-bit gray frames (CV_8UC1)")
        return np.ascontiguousarray(a)
