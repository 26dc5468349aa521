"""Deterministic synthetic (image, pose, depth) sequences.

Stand-in for the reference's external test data (``test/dataset.cpp:81-186``
reads ``first_200_frames_traj_over_table``; it is not in this repo and there is
no network).  Camera intrinsics follow ``test/dataset_main.cpp:37`` scaled to
the image size; see ``rmd_synth.c`` for the scene and trajectory.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librmd_synth.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile rmd_synth.c in-tree (host C + OpenMP)."""
    src = os.path.join(_HERE, "rmd_synth.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-O2", "-std=gnu11", "-fPIC", "-fopenmp", "-shared",
                           "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.rmd_synth_create.restype = ctypes.c_void_p
        lib.rmd_synth_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                         ctypes.c_float, ctypes.c_float, ctypes.c_uint32]
        lib.rmd_synth_destroy.argtypes = [ctypes.c_void_p]
        lib.rmd_synth_pose.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.rmd_synth_render.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 4
        _lib = lib
    return _lib


def dataset_camera(width: int, height: int):
    """fx, fy, cx, cy of the reference data set (test/dataset_main.cpp:37:
    481.2, -480, 319.5, 239.5 at 640x480), scaled to (width, height)."""
    return (481.2 * width / 640.0, -480.0 * height / 480.0, (width - 1) / 2.0, (height - 1) / 2.0)


def se3_inv(T: np.ndarray) -> np.ndarray:
    """Inverse of a 3x4 [R|t] in float32 (host-side helper)."""
    T = np.asarray(T, dtype=np.float32).reshape(3, 4)
    R = T[:, :3]
    out = np.empty((3, 4), dtype=np.float32)
    out[:, :3] = R.T
    out[:, 3] = -(R.T @ T[:, 3])
    return out


@dataclass
class Frame:
    index: int
    image_u8: np.ndarray      # (h, w) uint8
    image: np.ndarray         # (h, w) float32 = u8 * (1/255)
    depth: np.ndarray         # (h, w) float32, distance along the viewing ray
    T_world_cam: np.ndarray   # (3, 4) float32

    @property
    def T_cam_world(self) -> np.ndarray:
        """World -> camera: what rmd::SeedMatrix calls T_curr_world."""
        return se3_inv(self.T_world_cam)


class SyntheticSequence:
    """Frames are rendered lazily and deterministically from (size, seed)."""

    def __init__(self, width=640, height=480, seed=0x5EED0002, camera=None):
        self.width, self.height, self.seed = int(width), int(height), int(seed)
        self.fx, self.fy, self.cx, self.cy = camera if camera is not None else dataset_camera(width, height)
        self._lib = _load()
        self._h = self._lib.rmd_synth_create(self.width, self.height, self.fx, self.fy, self.cx, self.cy,
                                             self.seed & 0xFFFFFFFF)
        if not self._h:
            raise MemoryError("rmd_synth_create failed")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.rmd_synth_destroy(h)

    @property
    def camera(self):
        return (self.fx, self.fy, self.cx, self.cy)

    def pose(self, k: int) -> np.ndarray:
        T = np.empty((3, 4), dtype=np.float32)
        self._lib.rmd_synth_pose(self._h, int(k), T.ctypes.data)
        return T

    def frame(self, k: int, want_depth: bool = True) -> Frame:
        T = self.pose(k)
        u8 = np.empty((self.height, self.width), dtype=np.uint8)
        f32 = np.empty((self.height, self.width), dtype=np.float32)
        depth = np.empty((self.height, self.width), dtype=np.float32) if want_depth else None
        self._lib.rmd_synth_render(self._h, T.ctypes.data, u8.ctypes.data, f32.ctypes.data,
                                   depth.ctypes.data if want_depth else None)
        return Frame(k, u8, f32, depth, T)

    def frames(self, n: int, want_depth: bool = False):
        for k in range(n):
            yield self.frame(k, want_depth=want_depth or k == 0)
