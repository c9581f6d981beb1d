"""Host-only pieces of the C++ shim (tests/cpp/host_test.cpp): setFromTriplets semantics of sp_from_triplets, the CSV /
MeshLoader readers of mesh_io.h and the operator expression lowering -- none of them touches the device, so they run in
the CPU suite."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_side(fdb, tmp_path):
    exe = str(tmp_path / "host_test")
    lib = os.path.join(ROOT, "fdapde-core_b200", "lib")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_test.cpp"), "-L", lib, "-lfdapde_b200",
                           "-Wl,-rpath," + lib, "-o", exe])
    env = dict(os.environ, TMPDIR=str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
    print(r.stdout)
    assert r.returncode == 0 and "HOST_TEST_PASS" in r.stdout
