"""Frame ingest with lens undistortion -- SURVEY.md 8f row 1:
rmd::Depthmap::initUndistortionMap + inputImage (src/depthmap.cpp:45-61,95-106), i.e. OpenCV's
initUndistortRectifyMap (CV_16SC2 maps) -> remap(INTER_LINEAR, 8-bit) -> convertTo(CV_32F, 1/255.f).

Integer / fixed-point work: the bar is BIT-EXACT.
  * CPU: oracle/rmd_oracle_ingest.c against golden vectors made with the real OpenCV
    (tests/golden/make_golden_undistort.py; three cameras, one of them the reference's
    launch/px4_2.launch at full 752x480 through SHA-256 digests + a crop).
  * GPU: the product (C-ABI rmd_seeds_init_undistortion_map / undistort_u8 / *_u8) against the same
    vectors and against the oracle at VGA and 1080p; the filter fed distorted 8-bit frames equals the
    filter fed the oracle-undistorted float frames, bit for bit.
"""
import hashlib
import os

import numpy as np
import pytest

import oracle_binding as ob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "undistort_cv2.npz")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def texture(w, h, seed):
    """tests/golden/make_golden_undistort.py::texture (integer arithmetic only)."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    hsh = (xx * 73856093) ^ (yy * 19349663) ^ (seed * 83492791)
    hsh = (hsh ^ (hsh >> 13)) * 1274126177 & 0xFFFFFFFF
    noise = (hsh >> 16) % 61
    bands = ((xx * 7 + yy * 3) % 256 + (xx * yy // 17) % 128) // 2
    return np.clip(bands + noise, 0, 255).astype(np.uint8)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _cam(gold, name):
    w, h = (int(v) for v in gold[f"{name}_size"])
    return w, h, [float(v) for v in gold[f"{name}_camera"]], [float(v) for v in gold[f"{name}_dist"]]


# ------------------------------------------------------------------ CPU: the oracle is pinned by OpenCV
@pytest.mark.parametrize("name", ["small", "tilt"])
def test_oracle_matches_opencv_full_fixture(gold, name):
    w, h, cam, dist = _cam(gold, name)
    m1, m2 = ob.undistort_maps(w, h, *cam, *dist)
    assert np.array_equal(m1, gold[f"{name}_map1"]) and np.array_equal(m2, gold[f"{name}_map2"])
    und = ob.remap_u8(gold[f"{name}_img"], m1, m2)
    assert np.array_equal(und, gold[f"{name}_undistorted"])
    assert np.array_equal(ob.u8_to_float(und), gold[f"{name}_float"])
    if name == "tilt":   # the border-constant path is exercised
        assert int(gold["tilt_outside_pixels"]) > 1000 and (und == 0).sum() > 100


def test_oracle_matches_opencv_px4_launch_file_camera(gold):
    """launch/px4_2.launch at its full 752x480: digests of every product of the chain + a crop."""
    w, h, cam, dist = _cam(gold, "px4")
    assert (w, h) == (752, 480)
    m1, m2 = ob.undistort_maps(w, h, *cam, *dist)
    img = texture(w, h, int(gold["px4_img_seed"]))
    und = ob.remap_u8(img, m1, m2)
    flt = ob.u8_to_float(und)
    for key, arr in (("map1", m1), ("map2", m2), ("undistorted", und), ("float", flt)):
        assert np.array_equal(arr[:48, :64], gold[f"px4_{key}_crop"]), key
        assert sha(arr) == str(gold[f"px4_{key}_sha256"]), key


def test_identity_lens_and_float_conversion_properties():
    """k = 0: the maps are the pixel grid up to OpenCV's own rounding, remap is the identity away from it;
    convertTo(1/255.f) is monotone, exact at 0 and 255, and inverted by round(255 * f)."""
    w, h = 160, 120
    m1, m2 = ob.undistort_maps(w, h, 120.0, 119.0, 79.5, 59.5, 0.0, 0.0, 0.0, 0.0)
    yy, xx = np.mgrid[0:h, 0:w]
    exact = (m2 == 0) & (m1[..., 0] == xx) & (m1[..., 1] == yy)
    near = (np.abs(m1[..., 0] + (m2 & 31) / 32.0 - xx) <= 1 / 32 + 1e-9) & (np.abs(m1[..., 1] + (m2 >> 5) / 32.0 - yy) <= 1 / 32 + 1e-9)
    assert near.all() and exact.mean() > 0.5
    img = texture(w, h, 7)
    und = ob.remap_u8(img, m1, m2)
    assert np.array_equal(und[exact], img[exact])
    ramp = np.arange(256, dtype=np.uint8).reshape(16, 16)
    f = ob.u8_to_float(ramp)
    assert f[0, 0] == 0.0 and f[15, 15] == np.float32(255) * (np.float32(1) / np.float32(255))
    assert np.all(np.diff(f.ravel()) > 0) and np.array_equal(np.rint(f * 255).astype(np.uint8), ramp)


def test_oracle_matches_live_opencv_on_random_cameras():
    """Beyond the committed fixtures: wherever cv2 is importable (it is in this image), 40 seeded random
    cameras / lenses / sizes, maps + remap + conversion, all bit-exact."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(20260923)
    for case in range(40):
        w, h = int(rng.integers(24, 200)), int(rng.integers(24, 160))
        fx, fy = float(rng.uniform(0.5, 2.0) * w), float(rng.choice([-1, 1]) * rng.uniform(0.5, 2.0) * w)
        cx, cy = float(rng.uniform(0.3, 0.7) * w), float(rng.uniform(0.3, 0.7) * h)
        dist = [float(rng.uniform(-0.4, 0.4)), float(rng.uniform(-0.15, 0.15)), float(rng.uniform(-0.01, 0.01)),
                float(rng.uniform(-0.01, 0.01))]
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
        D = np.array([dist], np.float32)
        m1, m2 = cv2.initUndistortRectifyMap(K, D, np.eye(3), K, (w, h), cv2.CV_16SC2)
        cam = [float(v) for v in (K[0, 0], K[1, 1], K[0, 2], K[1, 2])]
        o1, o2 = ob.undistort_maps(w, h, *cam, *[float(v) for v in D[0]])
        assert np.array_equal(o1, m1) and np.array_equal(o2, m2), f"case {case}: maps differ"
        img = texture(w, h, case)
        und = cv2.remap(img, m1, m2, cv2.INTER_LINEAR)
        assert np.array_equal(ob.remap_u8(img, o1, o2), und), f"case {case}: remap differs"
        flt = cv2.multiply(und, 1.0, scale=float(np.float32(1.0) / np.float32(255.0)), dtype=cv2.CV_32F)
        assert np.array_equal(ob.u8_to_float(und), flt), f"case {case}: conversion differs"


def _product_maps_fn():
    """The product's own host code (csrc/undistort_maps.h), built into a shim with the system compiler the same way
    nvcc hands it to the host compiler (-O3, no fast-math)."""
    import ctypes
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "tests", "cpp", "build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "libundistort_maps_shim.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++14", "-O3", "-fPIC", "-shared", os.path.join(root, "tests", "cpp", "undistort_maps_shim.cpp"),
                           "-o", lib])
    L = ctypes.CDLL(lib)
    L.product_undistort_maps.argtypes = [ctypes.c_int] * 2 + [ctypes.c_float] * 8 + [ctypes.c_void_p] * 2

    def maps(w, h, fx, fy, cx, cy, k1, k2, p1, p2):
        m1, m2 = np.empty((h, w, 2), np.int16), np.empty((h, w), np.uint16)
        L.product_undistort_maps(w, h, fx, fy, cx, cy, k1, k2, p1, p2, m1.ctypes.data, m2.ctypes.data)
        return m1, m2
    return maps


def test_product_host_map_code_matches_opencv(gold):
    """Not only the oracle: the product's host-side map computation itself, on the CPU, against the golden vectors
    (three cameras incl. launch/px4_2.launch at 752x480) and, where cv2 is importable, 40 random cameras."""
    maps = _product_maps_fn()
    for name in ("small", "tilt"):
        w, h, cam, dist = _cam(gold, name)
        m1, m2 = maps(w, h, *cam, *dist)
        assert np.array_equal(m1, gold[f"{name}_map1"]) and np.array_equal(m2, gold[f"{name}_map2"]), name
    w, h, cam, dist = _cam(gold, "px4")
    m1, m2 = maps(w, h, *cam, *dist)
    assert sha(m1) == str(gold["px4_map1_sha256"]) and sha(m2) == str(gold["px4_map2_sha256"])
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for case in range(40):
        w, h = int(rng.integers(24, 260)), int(rng.integers(24, 200))
        K = np.array([[rng.uniform(0.5, 2.0) * w, 0, rng.uniform(0.3, 0.7) * w],
                      [0, rng.choice([-1, 1]) * rng.uniform(0.5, 2.0) * w, rng.uniform(0.3, 0.7) * h], [0, 0, 1]], np.float32)
        D = np.array([[rng.uniform(-0.4, 0.4), rng.uniform(-0.15, 0.15), rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01)]], np.float32)
        c1, c2 = cv2.initUndistortRectifyMap(K, D, np.eye(3), K, (w, h), cv2.CV_16SC2)
        m1, m2 = maps(w, h, float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), *[float(v) for v in D[0]])
        assert np.array_equal(m1, c1) and np.array_equal(m2, c2), f"case {case}"


# ------------------------------------------------------------------ GPU: the product
gpu = pytest.mark.gpu


def _handle(w, h, cam, dist=None):
    import rpg_open_remode_b200 as rmd
    g = rmd.SeedMatrix(w, h, rmd.PinholeCamera(*cam))
    if dist is not None:
        g.initUndistortionMap(*dist)
    return g


@gpu
@pytest.mark.parametrize("name", ["small", "tilt", "px4"])
def test_product_matches_opencv_golden(gold, name):
    import rpg_open_remode_b200 as rmd
    w, h, cam, dist = _cam(gold, name)
    g = _handle(w, h, cam, dist)
    m1, m2 = g.getUndistortionMap()
    img = gold[f"{name}_img"] if name != "px4" else texture(w, h, int(gold["px4_img_seed"]))
    und = g.undistort(img)
    # the float image the filter sees: set it as the reference frame and read it back
    g.setReferenceImage(img, np.eye(4, dtype=np.float32)[:3], 0.5, 2.0)
    flt = g._download(rmd.FIELD_REF_IMG)
    if name == "px4":
        for key, arr in (("map1", m1), ("map2", m2), ("undistorted", und), ("float", flt)):
            assert np.array_equal(arr[:48, :64], gold[f"px4_{key}_crop"]), key
            assert sha(arr) == str(gold[f"px4_{key}_sha256"]), key
    else:
        assert np.array_equal(m1, gold[f"{name}_map1"]) and np.array_equal(m2, gold[f"{name}_map2"])
        assert np.array_equal(und, gold[f"{name}_undistorted"])
        assert np.array_equal(flt, gold[f"{name}_float"])


@gpu
@pytest.mark.parametrize("size", [(640, 480), (1920, 1080), (101, 77)])
def test_product_matches_oracle_at_full_sizes(size):
    w, h = size
    cam = (481.2 * w / 640, -480.0 * h / 480, (w - 1) / 2, (h - 1) / 2)
    dist = (0.23, -0.11, 0.004, -0.003)      # pincushion: corners sample outside the source
    g = _handle(w, h, cam, dist)
    m1, m2 = g.getUndistortionMap()
    o1, o2 = ob.undistort_maps(w, h, *[float(np.float32(c)) for c in cam], *dist)
    assert np.array_equal(m1, o1) and np.array_equal(m2, o2)
    img = texture(w, h, 11)
    want = ob.remap_u8(img, o1, o2)
    assert np.array_equal(g.undistort(img), want)
    assert (want == 0).sum() > 0.001 * w * h     # the border path is exercised
    # idempotence of the maps, and switching the lens off again
    g.initUndistortionMap(*dist)
    assert np.array_equal(g.undistort(img), want)
    g.clearUndistortionMap()
    with pytest.raises(Exception):
        g.undistort(img)


@gpu
def test_filter_on_distorted_u8_frames_equals_filter_on_undistorted_floats(qvga_sequence):
    """The fused ingest (remap + 1/255) in front of the depth filter changes nothing else: 12 updates from
    'distorted' 8-bit frames == 12 updates from the oracle's undistorted float frames, bit for bit."""
    import rpg_open_remode_b200 as rmd
    seq = qvga_sequence
    w, h = seq.width, seq.height
    dist = (-0.21, 0.06, 0.001, -0.0005)
    cam = [float(np.float32(c)) for c in seq.camera]
    o1, o2 = ob.undistort_maps(w, h, *cam, *dist)
    a = _handle(w, h, seq.camera, dist)   # u8 in, undistorted on the GPU
    b = _handle(w, h, seq.camera)         # float in, undistorted by the oracle
    f0 = seq.frame(0)
    dmin, dmax = float(f0.depth.min()), float(f0.depth.max())
    a.setReferenceImage(f0.image_u8, f0.T_cam_world, dmin, dmax)
    b.setReferenceImage(ob.u8_to_float(ob.remap_u8(f0.image_u8, o1, o2)), f0.T_cam_world, dmin, dmax)
    for k in range(1, 13):
        f = seq.frame(k, want_depth=False)
        a.update(f.image_u8, f.T_cam_world)
        b.update(ob.u8_to_float(ob.remap_u8(f.image_u8, o1, o2)), f.T_cam_world)
    for get in ("downloadConvergence", "downloadDepthmap", "downloadSigmaSq", "downloadA", "downloadB"):
        assert np.array_equal(getattr(a, get)(), getattr(b, get)()), get
    # Depthmap facade: initUndistortionMap + 8-bit frames (src/depthmap.cpp:45-61,66-106)
    d = rmd.Depthmap(w, h, seq.camera[0], seq.camera[2], seq.camera[1], seq.camera[3])
    d.initUndistortionMap(*dist)
    d.setReferenceImage(f0.image_u8, f0.T_cam_world, dmin, dmax)
    assert np.array_equal(d.getReferenceImage(), ob.remap_u8(f0.image_u8, o1, o2))
