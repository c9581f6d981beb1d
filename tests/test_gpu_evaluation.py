"""GPU parity for the basis-evaluation row: point location, pointwise and areal Psi through the C ABI,
against the CPU oracle on the same inputs and against the reference's own .mtx fixtures
(test/src/lagrangian_basis_test.cpp:200-238; tolerance of the reference test: 1e-7, ours: 1e-12)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
PSI = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "psi.npz"))


def _space(fdb, pts, els, bnd, R):
    mesh = fdb.Triangulation(pts, els, bnd)
    basis = fdb.LagrangianBasis(mesh, R)
    return fdb.Space(mesh, R, basis.dofs(), basis.size(), basis.boundary_dofs()), basis


def _dense(shape, r, c, v):
    return sp.coo_matrix((v, (r, c)), shape=tuple(int(x) for x in shape)).toarray()


@pytest.mark.parametrize("R", [1, 2])
def test_pointwise_psi_matches_oracle_and_reference_fixture(fdb, golden_meshes, R):
    pts, els, bnd = golden_meshes("c_shaped")
    s, basis = _space(fdb, pts, els, bnd, R)
    locs = PSI["c_shaped/locs"]
    ids, cols, vals = s.eval_pointwise(locs)
    ids_o, cols_o, vals_o = orc.eval_pointwise(R, pts, els, basis.dofs(), locs)
    assert np.array_equal(ids, ids_o) and np.array_equal(cols, cols_o)     # indices bit-exact
    assert np.max(np.abs(vals - vals_o)) < 1e-12
    name = f"lagrangian_pointwise_eval_order{R}"
    rows = np.repeat(np.arange(locs.shape[0]), cols.shape[1])
    got = _dense((locs.shape[0], basis.size()), rows, cols.ravel(), vals.ravel())
    want = _dense(PSI[name + "/shape"], PSI[name + "/rows"], PSI[name + "/cols"], PSI[name + "/vals"])
    assert np.max(np.abs(got - want)) < 1e-12
    assert np.array_equal(s.locate(locs), ids)


@pytest.mark.parametrize("R", [1, 2])
def test_areal_psi_matches_oracle_and_reference_fixture(fdb, golden_meshes, R):
    pts, els, bnd = golden_meshes("quasi_circle")
    s, basis = _space(fdb, pts, els, bnd, R)
    inc = PSI["quasi_circle/incidence"]
    rows, cols, vals, D = s.eval_areal(inc)
    ro, co, vo, Do = orc.eval_areal(R, pts, els, basis.dofs(), inc)
    assert np.array_equal(rows, ro) and np.array_equal(cols, co)
    assert np.max(np.abs(vals - vo) / np.maximum(np.abs(vo), 1e-3)) < 1e-12
    assert np.max(np.abs(D - Do) / Do) < 1e-14
    name = f"lagrangian_areal_eval_order{R}"
    got = _dense((inc.shape[0], basis.size()), rows, cols, vals)
    want = _dense(PSI[name + "/shape"], PSI[name + "/rows"], PSI[name + "/cols"], PSI[name + "/vals"])
    assert np.max(np.abs(got - want)) < 1e-9   # the fixture itself carries the 15-digit quadrature table
    assert np.allclose(got.sum(axis=1), 1.0, atol=1e-12)


@pytest.mark.parametrize("dim", [2, 3])
def test_locate_random_boundary_and_outside_points(fdb, golden_meshes, dim):
    rng = np.random.default_rng(7)
    if dim == 2:
        pts, els, bnd = golden_meshes("unit_square")
    else:
        pts, els, bnd = golden_meshes("unit_sphere")
    s, _ = _space(fdb, pts, els, bnd, 1)
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    inside = lo + (hi - lo) * rng.random((4000, dim))
    # mesh nodes (shared by many cells), edge midpoints, barycentres and far-away points
    mid = 0.5 * (pts[els[:200, 0]] + pts[els[:200, 1]])
    bary = pts[els[:300]].mean(axis=1)
    far = hi + 1.0 + rng.random((50, dim))
    locs = np.concatenate([inside, pts[:400], mid, bary, far])
    ids = s.locate(locs)
    ids_o = orc.locate(pts, els, locs)
    assert np.array_equal(ids, ids_o)
    assert (ids[-50:] == -1).all() and (ids[-350:-50] == np.arange(300)).all()
    # a second query on the same space reuses the locator
    assert np.array_equal(s.locate(locs[::7]), ids[::7])


def test_pointwise_psi_large_structured_mesh_properties(fdb):
    """1.3 M triangles, 2 M points: partition of unity, reproduction of linear functions, every point located."""
    nodes, cells, bnd = fdb.meshes.unit_square(800)
    s, basis = _space(fdb, nodes, cells, bnd, 2)
    rng = np.random.default_rng(1)
    locs = rng.random((2_000_000, 2))
    ids, cols, vals = s.eval_pointwise(locs)
    assert (ids >= 0).all() and (cols >= 0).all()
    assert np.max(np.abs(vals.sum(axis=1) - 1.0)) < 1e-12
    xc = s.dofs_coords()
    lin = 3.0 * xc[:, 0] - 2.0 * xc[:, 1] + 0.5
    assert np.max(np.abs((vals * lin[cols]).sum(axis=1) - (3.0 * locs[:, 0] - 2.0 * locs[:, 1] + 0.5))) < 1e-11
    # the located cell really contains the point: barycentric coordinates of the P1 part are non-negative
    v = nodes[cells[ids]]
    T = np.stack([v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]], axis=2)
    z = np.linalg.solve(T, (locs - v[:, 0])[:, :, None])[:, :, 0]
    assert z.min() > -1e-12 and (1 - z.sum(axis=1)).min() > -1e-12


@pytest.mark.parametrize("R", [1, 2])
def test_surface_location_and_pointwise_psi(fdb, golden_meshes, R):
    pts, els, bnd = golden_meshes("surface")
    s, basis = _space(fdb, pts, els, bnd, R)
    rng = np.random.default_rng(5)
    ids0 = rng.integers(0, els.shape[0], 500)
    w = rng.dirichlet(np.ones(3), 500)
    p = np.einsum("ik,ikd->id", w, pts[els[ids0]])
    locs = np.concatenate([p, p[:20] + np.array([0.0, 0.0, 1e-9])])
    ids, cols, vals = s.eval_pointwise(locs)
    ids_o, cols_o, vals_o = orc.eval_pointwise(R, pts, els, basis.dofs(), locs)
    assert np.array_equal(ids, ids_o) and np.array_equal(cols, cols_o)
    assert np.max(np.abs(vals - vals_o)) < 1e-12
    assert np.array_equal(ids[:500], ids0) and (ids[500:] == -1).all()
