"""Builds libfdapde_b200.so in-tree with nvcc for sm_100a (run by __graft_entry__.build()).

    python fdapde-core_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfdapde_b200.so")
SOURCES = ["api.cu", "tables.cu", "pattern.cu", "assemble.cu", "solve.cu", "topology.cu", "comm.cu", "solve_persistent.cu", "solve_peer.cu", "evaluate.cu", "surface.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "local_matrix.cuh"), os.path.join(CSRC, "solve_common.cuh"), os.path.join(HERE, "..", "include", "fdapde_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), variant=None):
    """variant/defines: development builds of kernel variants (lib/libfdapde_b200_<variant>.so, selected at run time
    with FDB_LIB_PATH); the product library is the default build."""
    obj_dir = OBJ if variant is None else OBJ + "_" + variant
    lib = LIB if variant is None else os.path.join(LIBDIR, f"libfdapde_b200_{variant}.so")
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + HEADERS):
            cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=min(6, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(lib, objs):
        run([NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])
    return lib


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    var = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--variant=")), None)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, variant=var))
