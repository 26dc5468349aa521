"""Data-set and wire formats: rmd::test::Dataset (test/dataset.cpp:62-213), the DenseInput
field mapping (src/depthmap_node.cpp:97-132, test/publish_dataset.cpp:77-100) and the dataset_main protocol
(test/dataset_main.cpp:49-140) -- host logic, CPU only, on files written here in the data set's layout."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import pytest

from rpg_open_remode_b200 import dataset as ds
from rpg_open_remode_b200.api import SE3


def _write_png_gray(path, img):
    """Minimal 8-bit grayscale PNG writer (so the test does not depend on the reader's own library)."""
    import struct
    import zlib
    h, w = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


@pytest.fixture()
def tiny_dataset(tmp_path):
    w, h, n = 12, 8, 6
    os.makedirs(tmp_path / "images")
    os.makedirs(tmp_path / "depthmaps")
    rng = np.random.default_rng(5)
    imgs, depths, lines = [], [], []
    for k in range(n):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        depth_cm = rng.uniform(80, 300, (h, w)).astype(np.float32)
        name = f"scene_{k:03d}.png"
        _write_png_gray(str(tmp_path / "images" / name), img)
        with open(tmp_path / "depthmaps" / f"scene_{k:03d}.depth", "w") as f:
            f.write(" ".join(f"{v:.4f}" for v in depth_cm.ravel()) + "\n")
        lines.append(f"{name} {0.1 * k:.3f} {-0.2 * k:.3f} 1.5 0.0 0.3826834 0.0 0.9238795")
        imgs.append(img)
        depths.append(depth_cm)
    (tmp_path / "seq.txt").write_text("\n".join(lines) + "\n")
    return str(tmp_path), imgs, depths, (w, h, n)


def test_sequence_file_windowing_and_fields(tiny_dataset):
    path, imgs, depths, (w, h, n) = tiny_dataset
    d = ds.Dataset("seq.txt", path)
    assert d.readDataSequence() and len(d) == n                     # (0, 0) = everything, test/dataset.cpp:135-138
    assert d.readDataSequence(2, 5) and [e.image_file_name for e in d] == [f"scene_{k:03d}.png" for k in (2, 3, 4)]
    assert d.readDataSequence(4, 0) and len(d) == 2                 # end == 0: to the end of the file (:97)
    assert d.readDataSequence(3, 2) and len(d) == 0
    d.readDataSequence()
    e = d(1)
    assert e.depthmap_file_name == "scene_001.depth"                # stem up to the FIRST dot + "depth" (:106)
    assert np.allclose(e.translation, [0.1, -0.2, 1.5]) and np.allclose(e.quaternion, [0, 0.3826834, 0, 0.9238795])
    with pytest.raises(IndexError):
        d(n)
    assert not ds.Dataset().readDataSequence() and not ds.Dataset("missing.txt", path).readDataSequence()


def test_image_depth_and_pose_readers(tiny_dataset):
    path, imgs, depths, (w, h, n) = tiny_dataset
    d = ds.Dataset("seq.txt", path)
    d.readDataSequence()
    for k in (0, 3):
        e = d(k)
        assert np.array_equal(d.readImage(e), imgs[k])
        got = d.readDepthmap(e, w, h)
        assert got.dtype == np.float32 and np.allclose(got, depths[k] / 100.0, rtol=0, atol=1e-6)   # cm -> m (:185)
    assert d.readImage("nope.png") is None and d.readDepthmap(ds.DatasetEntry(depthmap_file_name="nope.depth"), w, h) is None
    # pose: the file stores qx qy qz qw, SE3 takes (qw, qx, qy, qz, t) -- a 45 degree turn about y
    T = ds.Dataset.readCameraPose(d(2))
    assert np.allclose(T.data.reshape(3, 4)[:, 3], [0.2, -0.4, 1.5])
    c = np.float32(np.sqrt(0.5))
    assert np.allclose(T.data.reshape(3, 4)[:, :3], [[c, 0, c], [0, 1, 0], [-c, 0, c]], atol=1e-6)
    assert np.array_equal(T.data, SE3(0.9238795, 0.0, 0.3826834, 0.0, 0.2, -0.4, 1.5).data)


def test_env_var_and_dense_input_mapping(tiny_dataset, monkeypatch):
    path, imgs, depths, (w, h, n) = tiny_dataset
    d = ds.Dataset("seq.txt")
    monkeypatch.delenv(ds.DATA_PATH_ENV_VAR, raising=False)
    assert not d.loadPathFromEnv() and not d.readDataSequence()
    monkeypatch.setenv(ds.DATA_PATH_ENV_VAR, path)
    assert ds.DATA_PATH_ENV_VAR == "RMD_TEST_DATA_PATH" and d.loadPathFromEnv() and d.readDataSequence()
    msg = ds.DenseInput.from_dataset(d, d(4), frame_id=4)
    assert np.array_equal(msg.image, imgs[4]) and msg.frame_id == 4
    assert msg.orientation_wxyz == pytest.approx((0.9238795, 0.0, 0.3826834, 0.0)) and msg.position_xyz == pytest.approx((0.4, -0.8, 1.5))
    assert msg.min_depth == pytest.approx(depths[4].min() / 100, rel=1e-6) and msg.max_depth == pytest.approx(depths[4].max() / 100, rel=1e-6)
    assert np.array_equal(msg.T_world_curr().data, ds.Dataset.readCameraPose(d(4)).data)


def test_dataset_main_protocol_call_order(tiny_dataset):
    path, imgs, depths, (w, h, n) = tiny_dataset
    d = ds.Dataset("seq.txt", path)
    d.readDataSequence()
    os.remove(os.path.join(path, "images", "scene_003.png"))       # an unreadable frame is skipped (:63-67)
    calls, logged = [], []

    class FakeDepthmap:
        def setReferenceImage(self, img, T, dmin, dmax):
            calls.append(("ref", img.copy(), T.data.copy(), dmin, dmax))
            return True

        def update(self, img, T):
            calls.append(("update", img.copy(), T.data.copy()))

        def downloadDepthmap(self):
            calls.append(("download",))

        def downloadDenoisedDepthmap(self, lam, it):
            calls.append(("denoise", lam, it))

        def getDepthmap(self):
            return np.full((h, w), len(calls), np.float32)

    res = ds.run_dataset_experiment(FakeDepthmap(), d, w, h, log=logged.append)
    assert [c[0] for c in calls] == ["ref"] + ["update"] * 4 + ["download", "denoise"] and calls[-1] == ("denoise", 0.5, 200)
    assert res["updates"] == 4 and len(logged) == 1 and "scene_003.png" in logged[0]
    assert np.array_equal(calls[0][1], imgs[0]) and np.array_equal(calls[3][1], imgs[4])
    assert calls[0][3] == pytest.approx(depths[0].min() / 100, rel=1e-6)      # min/max of the frame's own depth map (:75-76)
    assert np.array_equal(calls[2][2], ds.Dataset.readCameraPose(d(2)).inv().data)   # T_curr_world = T_world_curr.inv() (:88,:101)
    assert res["depthmap"][0, 0] == 6 and res["denoised"][0, 0] == 7 and res["mean_update_s"] >= 0 and res["var_update_s"] >= 0


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_dataset_experiment_end_to_end_on_gpu(tmp_path):
    """test/dataset_main.cpp end to end on the GPU (tools/dataset_main.py's path): a sequence written in the data
    set's on-disk formats (sequence file, PNG, .depth text in cm) is read back by rmd::test::Dataset's mirror and
    driven through rmd::Depthmap (8-bit frames, poses parsed from quaternions) -- the result equals feeding the same
    parsed frames to a SeedMatrix by hand, converges, and reproduces the ray-cast ground truth."""
    import importlib.util
    import rpg_open_remode_b200 as rmd
    from rpg_open_remode_b200 import dataset as ds
    spec = importlib.util.spec_from_file_location("dataset_main", os.path.join(ROOT, "tools", "dataset_main.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    W, H, N = 320, 240, 60
    cam = tool.write_synthetic(str(tmp_path), n=N, w=W, h=H, seed=0x5EED0002)
    d = ds.Dataset("first_200_frames_traj_over_table_input_sequence.txt", str(tmp_path))
    assert d.readDataSequence(0, 200) and len(d) == N
    depthmap = rmd.Depthmap(W, H, cam[0], cam[2], cam[1], cam[3])
    res = ds.run_dataset_experiment(depthmap, d, W, H, log=lambda *a: None)
    assert res["updates"] == N - 1 and res["depthmap"].shape == (H, W)
    # by hand, through the SeedMatrix-level API, with the same parsed inputs
    seeds = rmd.SeedMatrix(W, H, rmd.PinholeCamera(*cam))
    gt = None
    for k, e in enumerate(d):
        img = d.readImage(e)
        T = ds.Dataset.readCameraPose(e).inv()
        if k == 0:
            gt = d.readDepthmap(e, W, H)
            seeds.setReferenceImage(img, T, float(gt.min()), float(gt.max()))
        else:
            seeds.update(img, T)
    assert np.array_equal(res["depthmap"], seeds.downloadDepthmap())
    depthmap.downloadConvergenceMap()
    conv = depthmap.getConvergenceMap()
    assert np.array_equal(conv, seeds.downloadConvergence())
    c = conv == 1
    rng_d = float(gt.max() - gt.min())
    assert c.mean() > 0.5 and depthmap.getConvergedPercentage() == pytest.approx(100.0 * c.mean(), abs=1e-3)
    assert np.median(np.abs(res["depthmap"] - gt)[c]) < 0.01 * rng_d
    # the denoised map (0.5, 200 as test/dataset_main.cpp:116) stays inside the raw map's range and close to it
    den = res["denoised"]
    assert den.min() >= res["depthmap"].min() - 1e-5 and den.max() <= res["depthmap"].max() + 1e-5
    assert np.median(np.abs(den - gt)[c]) < 0.01 * rng_d
