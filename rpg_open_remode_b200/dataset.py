"""Data-set and wire formats either side of the path, host side only.

  Dataset      rmd::test::Dataset (test/dataset.cpp:62-213): the sequence file of the REMODE data sets
               (`<image file> tx ty tz qx qy qz qw` per line), 8-bit images under images/, ground-truth depth
               along the optical axis in centimetres as text under depthmaps/ (`<image stem>.depth`).
  DenseInput   svo_msgs/DenseInput as rmd::DepthmapNode consumes it (src/depthmap_node.cpp:97-132) and
               test/publish_dataset.cpp:77-100 fills it: MONO8 image, pose = orientation (w x y z) + position of
               the camera in the world, min / max scene depth.
  run_dataset_experiment   the protocol of test/dataset_main.cpp:32-140 (the reference's offline benchmark):
               first frame = reference, every further frame an update() that is timed, then the raw and the
               denoised (0.5, 200) depth maps.

Nothing numeric here: images and poses are handed to rpg_open_remode_b200.api unchanged.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass, field
from typing import Iterator, List, Optional

import numpy as np

from .api import SE3

DATA_PATH_ENV_VAR = "RMD_TEST_DATA_PATH"   # test/dataset.h:84


@dataclass
class DatasetEntry:   # test/dataset.h:33-54
    image_file_name: str = ""
    depthmap_file_name: str = ""
    translation: np.ndarray = field(default_factory=lambda: np.zeros(3, np.float32))   # x y z
    quaternion: np.ndarray = field(default_factory=lambda: np.array([0, 0, 0, 1], np.float32))   # x y z w (file order)


class Dataset:
    def __init__(self, sequence_file: str = "", dataset_path: str = ""):
        self.dataset_path_, self.sequence_file_ = dataset_path, sequence_file
        self.dataset_: List[DatasetEntry] = []

    def loadPathFromEnv(self) -> bool:                                   # test/dataset.cpp:199-208
        p = os.environ.get(DATA_PATH_ENV_VAR)
        if p is None:
            return False
        self.dataset_path_ = p
        return True

    def readDataSequence(self, start: int = 0, end: int = 0) -> bool:    # test/dataset.cpp:83-133
        if not self.dataset_path_ or not self.sequence_file_:
            return False
        self.dataset_ = []
        try:
            f = open(os.path.join(self.dataset_path_, self.sequence_file_))
        except OSError:
            return False
        with f:
            for line_cnt, line in enumerate(f):
                if line_cnt >= start and (line_cnt < end or end == 0):
                    tok = line.split()
                    e = DatasetEntry()
                    if tok:
                        e.image_file_name = tok[0]
                        # imgFileName.substr(0, imgFileName.find('.') + 1) + "depth"   (:106)
                        dot = tok[0].find(".")
                        e.depthmap_file_name = (tok[0][:dot + 1] if dot >= 0 else "") + "depth"
                        vals = [np.float32(v) for v in tok[1:8]]
                        vals += [np.float32(0)] * (7 - len(vals))   # a short line leaves the rest at its default
                        e.translation = np.array(vals[0:3], np.float32)
                        e.quaternion = np.array(vals[3:7], np.float32)
                    else:
                        e.depthmap_file_name = "depth"
                    self.dataset_.append(e)
        return True

    def readImage(self, entry) -> Optional[np.ndarray]:                   # :140-154: imread(..., GRAYSCALE)
        name = entry.image_file_name if isinstance(entry, DatasetEntry) else entry
        path = os.path.join(self.dataset_path_, "images", name)
        if not os.path.isfile(path):
            return None
        try:
            import cv2
            img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
            return None if img is None else np.ascontiguousarray(img, np.uint8)
        except ImportError:
            from PIL import Image
            return np.ascontiguousarray(Image.open(path).convert("L"), np.uint8)

    @staticmethod
    def readCameraPose(entry: DatasetEntry) -> SE3:                       # :156-167: SE3(qw, qx, qy, qz, tx, ty, tz)
        q, t = entry.quaternion, entry.translation
        return SE3(q[3], q[0], q[1], q[2], t[0], t[1], t[2])

    def readDepthmap(self, entry: DatasetEntry, width: int, height: int) -> Optional[np.ndarray]:   # :169-193
        path = os.path.join(self.dataset_path_, "depthmaps", entry.depthmap_file_name)
        try:
            with open(path) as f:
                vals = np.array(f.read().split()[:width * height], dtype=np.float32)
        except OSError:
            return None
        out = np.zeros(width * height, np.float32)
        out[:len(vals)] = vals
        return (out / np.float32(100.0)).reshape(height, width)           # centimetres -> metres (:185)

    def __iter__(self) -> Iterator[DatasetEntry]:
        return iter(self.dataset_)

    def __len__(self) -> int:
        return len(self.dataset_)

    def __call__(self, index: int) -> DatasetEntry:                       # :195-198 (operator(), bounds-checked)
        if not 0 <= index < len(self.dataset_):
            raise IndexError(index)
        return self.dataset_[index]


@dataclass
class DenseInput:
    """svo_msgs/DenseInput: what the node gets per frame (src/depthmap_node.cpp:97-132)."""
    image: np.ndarray            # MONO8
    orientation_wxyz: tuple
    position_xyz: tuple
    min_depth: float
    max_depth: float
    frame_id: int = 0

    def T_world_curr(self) -> SE3:                                        # src/depthmap_node.cpp:112-119
        w, x, y, z = self.orientation_wxyz
        return SE3(w, x, y, z, *self.position_xyz)

    @staticmethod
    def from_dataset(ds: Dataset, entry: DatasetEntry, frame_id: int = 0) -> Optional["DenseInput"]:
        """test/publish_dataset.cpp:55-100."""
        img = ds.readImage(entry)
        if img is None:
            return None
        depth = ds.readDepthmap(entry, img.shape[1], img.shape[0])
        if depth is None:
            return None
        q, t = entry.quaternion, entry.translation
        return DenseInput(img, (float(q[3]), float(q[0]), float(q[1]), float(q[2])),
                          (float(t[0]), float(t[1]), float(t[2])), float(depth.min()), float(depth.max()), frame_id)


def run_dataset_experiment(depthmap, dataset: Dataset, width: int = 640, height: int = 480, log=print) -> dict:
    """test/dataset_main.cpp:49-140 with the windows replaced by returned arrays.  `depthmap` is an
    rpg_open_remode_b200.api.Depthmap (or anything with its interface)."""
    first_img, update_time = True, []
    for data in dataset:
        img = dataset.readImage(data)
        if img is None:
            log(f"ERROR: could not read image {data.image_file_name}")          # :63-67
            continue
        depth = dataset.readDepthmap(data, img.shape[1], img.shape[0])
        if depth is None:
            log(f"ERROR: could not read depthmap {data.depthmap_file_name}")    # :70-74
            continue
        min_depth, max_depth = float(depth.min()), float(depth.max())          # :75-76
        T_world_curr = Dataset.readCameraPose(data)
        if first_img:
            if not depthmap.setReferenceImage(img, T_world_curr.inv(), min_depth, max_depth):   # :88
                raise RuntimeError("could not set reference image")
            first_img = False
        else:
            t0 = time.perf_counter()
            depthmap.update(img, T_world_curr.inv())                           # :100-104
            update_time.append(time.perf_counter() - t0)
    depthmap.downloadDepthmap()                                                # :109
    raw = np.array(depthmap.getDepthmap(), copy=True)
    depthmap.downloadDenoisedDepthmap(0.5, 200)                                # :115
    denoised = np.array(depthmap.getDepthmap(), copy=True)
    t = np.array(update_time, np.float64)
    mean = float(t.mean()) if len(t) else float("nan")
    var = float(((t - mean) ** 2).mean()) if len(t) else float("nan")          # :122-130
    return {"depthmap": raw, "denoised": denoised, "updates": len(t), "mean_update_s": mean, "var_update_s": var}
