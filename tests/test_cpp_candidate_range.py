"""CPU-runnable check of the candidate logic the kernels share (csrc/candidate_range.cuh): the header nvcc compiles
for the device is compiled here with the system compiler and compared, on 320 000 random and adversarial segments,
with the literal transcription of the reference's search loop (src/epipolar_match.cu:85-97): candidate count,
checkpoints, every restarted `l` bit for bit, and the exact first / last accepted candidate."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def test_candidate_range_against_the_reference_loop():
    out = os.path.join(ROOT, "tests", "cpp", "build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "candidate_range_test")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # -ffp-contract=off: the host must not fuse what the source does not fuse (the device code is written with explicit
    # fma_rn where it fuses)
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(CUDA, "include"), os.path.join(ROOT, "tests", "cpp", "candidate_range_test.cpp"),
                           "-o", exe])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "ALL CANDIDATE RANGE TESTS PASSED" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]
