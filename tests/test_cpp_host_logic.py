"""The ROS-free C++ mirrors of the path's callers: include/rmd/keyframe_node.h
(rmd::DepthmapNode's state machine) and include/rmd/dataset_reader.h (rmd::test::Dataset) are host-only headers;
tests/cpp/host_logic_test.cpp is compiled with the system compiler against include/ and run here (no GPU)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def test_keyframe_node_and_dataset_reader(tmp_path):
    out = os.path.join(ROOT, "tests", "cpp", "build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "host_logic_test")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(CUDA, "include"), os.path.join(ROOT, "tests", "cpp", "host_logic_test.cpp"),
                           "-o", exe])
    res = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "ALL HOST LOGIC TESTS PASSED" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
