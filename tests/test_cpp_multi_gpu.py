"""The C-ABI multi-GPU path (include/rmd_b200.h rmd_multi_*: ncclCommInitAll + grouped ncclSend/ncclRecv gather)
driven from plain C++ by one host thread: tests/cpp/multi_gpu_test.cpp compiles with the host compiler and, on the
GPU box, gathers one keyframe per visible GPU (1 GPU = a one-rank communicator) and checks the gathered maps
against each keyframe's own downloads.  The torch.distributed form of the same gather is covered on CPU with gloo
(tests/test_multi_gpu_host_logic.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "build", "multi_gpu_test")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build():
    from rpg_open_remode_b200 import _build as b, synth
    b.build_cuda()
    synth.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    pkg = os.path.join(ROOT, "rpg_open_remode_b200")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"),
           os.path.join(ROOT, "tests", "cpp", "multi_gpu_test.cpp"), "-o", EXE,
           "-L" + pkg, "-lrmd_b200", "-L" + os.path.join(pkg, "synth"), "-lrmd_synth",
           "-L" + os.path.join(CUDA, "lib64"), "-lcudart",
           "-Wl,-rpath," + pkg + ":" + os.path.join(pkg, "synth") + ":" + os.path.join(CUDA, "lib64")]
    subprocess.check_call(cmd)
    return EXE


def test_multi_gpu_cpp_caller_compiles():
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_multi_gpu_gather_from_cpp():
    exe = _build()
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(res.stdout[-2000:], res.stderr[-2000:])
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-2000:])
    assert "MULTI GPU TEST PASSED" in res.stdout
