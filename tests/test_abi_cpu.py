"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares, the
host-side operator algebra lowers expression trees like the reference, and compute calls fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            names |= set(re.findall(r"\b(fdb_[a-z0-9_]+)\s*\(", open(os.path.join(inc, f)).read()))
    return sorted(names)


def test_library_exports_every_declared_symbol(fdb):
    L = fdb.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ but not exported"
    assert L.fdb_version() >= 100


def test_operator_algebra_matches_reference_traits(fdb):
    # differential_expressions.h:70-73: symmetric = AND over leaves, Advection is not symmetric
    L = -fdb.laplacian() + fdb.advection([-1.0, 0.0]) + 2.0 * fdb.reaction(1.5)
    assert not L.is_symmetric
    assert [(k, s) for k, s, *_ in L.leaves] == [(0, -1.0), (2, 1.0), (3, 2.0)]
    assert (-fdb.laplacian() + fdb.reaction(1.0)).is_symmetric
    M = fdb.laplacian() - fdb.diffusion(np.eye(2))
    assert [(k, s) for k, s, *_ in M.leaves] == [(0, 1.0), (1, -1.0)]
    d = L.descriptor()
    assert d.n_terms == 3 and d.symmetric == 0 and d.terms[2].scale == 2.0
    assert fdb.dt().is_symmetric


def test_synthetic_meshes(fdb, golden_meshes):
    pts, els, bnd = golden_meshes("unit_square_16")
    n, c, b = fdb.meshes.unit_square(16)
    assert np.array_equal(c, els) and np.array_equal(b, bnd) and np.abs(n - pts).max() == 0.0
    n, c, b = fdb.meshes.unit_cube(4)
    assert n.shape == (125, 3) and c.shape == (384, 4) and int(b.sum()) == 125 - 27
    v = n[c]
    vol = np.abs(np.linalg.det(v[:, 1:] - v[:, :1])) / 6
    assert np.isclose(vol.sum(), 1.0) and np.allclose(vol, 1 / 384)


def test_compute_fails_loudly_without_gpu(fdb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    n, c, b = fdb.meshes.unit_square(4)
    with pytest.raises(fdb.FdbError) as ei:
        fdb.Assembler(fdb.Triangulation(n, c, b), 1, n.shape[0], c)
    assert ei.value.code == 2  # FDB_ERR_CUDA: no silent CPU fallback


def test_argument_errors_do_not_throw_across_abi(fdb):
    L = fdb.lib()
    h = C.c_void_p()
    assert L.fdb_space_create(C.byref(h), 1, 2, 1, 4, 2, None, None, 4, None) == 5  # network meshes: unsupported
    assert b"supported meshes" in L.fdb_last_error()
    assert L.fdb_space_create(C.byref(h), 2, 3, 1, 4, 2, None, None, 4, None) == 1  # surface meshes are accepted: null arrays
    assert L.fdb_space_create(C.byref(h), 2, 2, 1, 4, 2, None, None, 4, None) == 1  # null arrays
