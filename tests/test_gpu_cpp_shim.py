"""The C++ header shim (include/fdapde_b200/assembler.h) over the C ABI, exercised by a C++ program that mirrors the
reference's own fem_operators_test / fem_pde_test cases (tests/cpp/shim_test.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_shim_test():
    exe = os.path.join(ROOT, "tests", "cpp", "shim_test")
    src = os.path.join(ROOT, "tests", "cpp", "shim_test.cpp")
    lib = os.path.join(ROOT, "fdapde-core_b200", "lib")
    hdr = os.path.join(ROOT, "include", "fdapde_b200", "assembler.h")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-L", lib,
                               "-lfdapde_b200", "-Wl,-rpath," + lib, "-o", exe])
    return exe


def test_cpp_shim_compiles_and_fails_loudly_without_gpu(fdb):
    import torch
    exe = build_shim_test()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "no CUDA device" in r.stdout  # no silent CPU fallback


@pytest.mark.gpu
def test_cpp_shim_reference_cases(fdb):
    exe = build_shim_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout[-2000:])
    assert r.returncode == 0 and "SHIM_TEST_PASS" in r.stdout
