"""The header-compatible C++ classes in include/rmd/ (rmd::SeedMatrix, DepthmapDenoiser,
ImageReducer, DeviceImage, SE3, PinholeCamera, CudaException) compile with a plain
host compiler against the C-ABI and, on the GPU, pass the reference's gtests
re-hosted in tests/cpp/facade_test.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "build", "facade_test")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(patch=5):
    from rpg_open_remode_b200 import _build as b, synth
    b.build_cuda()
    synth.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    exe = EXE if patch == 5 else EXE + "_p%d" % patch
    pkg = os.path.join(ROOT, "rpg_open_remode_b200")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++14", "-O1", "-DRMD_BUILD_TESTS=1", "-DRMD_CORR_PATCH_SIDE=%d" % patch,
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"),
           os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"), "-o", exe,
           "-L" + pkg, "-lrmd_b200", "-L" + os.path.join(pkg, "synth"), "-lrmd_synth",
           "-L" + os.path.join(CUDA, "lib64"), "-lcudart",
           "-Wl,-rpath," + pkg + ":" + os.path.join(pkg, "synth") + ":" + os.path.join(CUDA, "lib64")]
    subprocess.check_call(cmd)
    return exe


def test_facade_headers_compile_with_host_compiler():
    exe = _build(5)
    assert os.path.exists(exe)


@pytest.mark.gpu
@pytest.mark.parametrize("patch", [5, 7])
def test_reference_gtests_rehosted_on_facade(patch):
    exe = _build(patch)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(res.stdout[-2000:], res.stderr[-2000:])
    assert res.returncode == 0, res.stdout[-2000:]
    assert "ALL FACADE TESTS PASSED" in res.stdout
