"""test/dataset_main.cpp on this library: the reference's offline experiment on the traj_over_table data set
(first 200 frames), when the data is available (RMD_TEST_DATA_PATH), else on the synthetic stand-in written to a
temporary directory in the same file formats.  GPU box.

    RMD_TEST_DATA_PATH=/data/traj_over_table python tools/dataset_main.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpg_open_remode_b200 as rmd  # noqa: E402
from rpg_open_remode_b200 import dataset as ds  # noqa: E402


def write_synthetic(path, n=200, w=640, h=480, seed=0x5EED0001):
    """A synthetic sequence in the data set's on-disk layout (sequence file, images/*.png, depthmaps/*.depth)."""
    import cv2
    from rpg_open_remode_b200 import synth
    seq = synth.SyntheticSequence(w, h, seed=seed)
    os.makedirs(os.path.join(path, "images"))
    os.makedirs(os.path.join(path, "depthmaps"))
    lines = []
    for k in range(n):
        f = seq.frame(k, want_depth=True)
        name = f"scene_{k:03d}.png"
        cv2.imwrite(os.path.join(path, "images", name), f.image_u8)
        np.savetxt(os.path.join(path, "depthmaps", name[:-3] + "depth"), (f.depth * 100.0).reshape(1, -1), fmt="%.4f")
        T_world_cam = rmd.SE3(f.T_cam_world).inv()
        R = T_world_cam.data.reshape(3, 4)[:, :3].astype(np.float64)
        t = T_world_cam.data.reshape(3, 4)[:, 3]
        qw = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
        qx, qy, qz = (R[2, 1] - R[1, 2]) / (4 * qw), (R[0, 2] - R[2, 0]) / (4 * qw), (R[1, 0] - R[0, 1]) / (4 * qw)
        lines.append(f"{name} {t[0]:.7f} {t[1]:.7f} {t[2]:.7f} {qx:.7f} {qy:.7f} {qz:.7f} {qw:.7f}")
    with open(os.path.join(path, "first_200_frames_traj_over_table_input_sequence.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return seq.camera


def main():
    d = ds.Dataset("first_200_frames_traj_over_table_input_sequence.txt")      # test/dataset_main.cpp:39
    cam = (481.2, -480.0, 319.5, 239.5)                                        # :37 (fx, fy, cx, cy)
    tmp = None
    if not d.loadPathFromEnv():
        print(f"'{ds.DATA_PATH_ENV_VAR}' is not set: writing a synthetic sequence in the data set's format")
        tmp = tempfile.TemporaryDirectory()
        cam = write_synthetic(tmp.name)
        d.dataset_path_ = tmp.name
    if not d.readDataSequence(0, 200):                                         # :45
        sys.exit("ERROR: could not read dataset")
    depthmap = rmd.Depthmap(640, 480, cam[0], cam[2], cam[1], cam[3])         # :55 (width, height, fx, cx, fy, cy)
    res = ds.run_dataset_experiment(depthmap, d)
    print(f"updates: {res['updates']}  MEAN update time: {res['mean_update_s']:.6f} s  "
          f"(STDDEV: {np.sqrt(res['var_update_s']):.6f})")
    print(f"converged: {depthmap.getConvergedPercentage():.1f} %   depth range of the result: "
          f"{res['depthmap'].min():.3f} .. {res['depthmap'].max():.3f} m")


if __name__ == "__main__":
    main()
