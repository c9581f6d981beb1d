"""Pins the CPU oracle (oracle/fdapde_oracle.c) against every known-answer test the reference holds for the
assembly + solve path.  CPU only.  All file:line citations are relative to /root/reference."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DOUBLE_TOLERANCE = 1e-7  # test/src/utils/constants.h:11


def almost_equal(a, b, eps=DOUBLE_TOLERANCE):  # test/src/utils/utils.h:32-41
    return abs(a - b) < eps or abs(a - b) < max(abs(a), abs(b)) * eps


def csc(outer, inner, val, n):
    return sp.csc_matrix((val, inner, outer), shape=(n, n))


# ---- reference element / basis -------------------------------------------------------------------------------

def test_poly_table_order():
    # multivariate_polynomial.h:52-79, first coordinate fastest
    assert orc.poly_table(2, 2).tolist() == [[0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [0, 2]]
    assert orc.poly_table(3, 1).tolist() == [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]
    assert orc.poly_table(3, 2).shape == (10, 3)


@pytest.mark.parametrize("M,R", [(1, 1), (1, 2), (2, 1), (2, 2), (3, 1), (3, 2)])
def test_lagrange_property(M, R):
    # lagrangian_basis_test.cpp:80-101: psi_i(node_j) = delta_ij
    c = orc.ref_basis_coeffs(M, R)
    nodes = orc.reference_nodes(M, R)
    for i in range(c.shape[0]):
        for j in range(c.shape[0]):
            assert almost_equal(orc.poly_eval(M, R, c[i], nodes[j]), 1.0 if i == j else 0.0)


def test_order1_reference_gradients():
    # lagrangian_basis_test.cpp:104-119
    c = orc.ref_basis_coeffs(2, 1)
    expected = [(-1.0, -1.0), (1.0, 0.0), (0.0, 1.0)]
    for i in range(3):
        for d in range(2):
            assert almost_equal(orc.poly_grad(2, 1, c[i], d, [0.0, 0.0]), expected[i][d])


def test_order2_reference_gradients():
    # lagrangian_basis_test.cpp:122-146
    c = orc.ref_basis_coeffs(2, 2)
    p = [0.5, 0.5]
    exp = [[1 - 4 * (1 - p[0] - p[1])] * 2, [4 * p[0] - 1, 0], [0, 4 * p[1] - 1],
           [4 * (1 - 2 * p[0] - p[1]), -4 * p[0]], [-4 * p[1], 4 * (1 - p[0] - 2 * p[1])], [4 * p[1], 4 * p[0]]]
    for i in range(6):
        for d in range(2):
            assert almost_equal(orc.poly_grad(2, 2, c[i], d, p), exp[i][d])


def test_order1_physical_gradients(golden_meshes):
    # lagrangian_basis_test.cpp:150-171, c_shaped cell 175, first node of the 6-point rule
    pts, els, _ = golden_meshes("c_shaped")
    p = orc.quadrature_table(2, 6)[0][0]
    g = orc.physical_gradients(2, 1, pts[els[175]], p)
    exp = np.array([[-5.2557081783567776, -3.5888585000106943], [6.2494499783110298, -1.2028086954015513],
                    [-0.9937417999542519, 4.7916671954122458]])
    assert np.max(np.abs(g - exp)) < 1e-12


def test_order2_physical_gradients(golden_meshes):
    # lagrangian_basis_test.cpp:174-197
    pts, els, _ = golden_meshes("c_shaped")
    p = orc.quadrature_table(2, 6)[0][0]
    g = orc.physical_gradients(2, 2, pts[els[175]], p)
    exp = np.array([[2.9830765115928704, 2.0369927574935405], [4.8982811692194259, -0.9427541948981384],
                    [-0.7788888242446018, 3.7556798236502558], [-6.6727629051483941, -6.9218931297696376],
                    [-9.8048064747509027, -4.3298093852388320], [9.3751005233316018, 6.4017841287628112]])
    assert np.max(np.abs(g - exp)) < 1e-12


# ---- local matrix ------------------------------------------------------------------------------------------------

GOLDEN_P2_STIFF = [  # fem_operators_test.cpp:83-96
    0.7043890316492852, 0.1653830261033185, 0.0694133177797771, -0.6615321044132733, -0.2776532711191089,
    0.0000000000000013, 0.1653830261033185, 0.7043890316492852, 0.0694133177797769, -0.6615321044132735,
    0.0000000000000003, -0.2776532711191076, 0.0694133177797771, 0.0694133177797769, 0.4164799066786617,
    0.0000000000000002, -0.2776532711191083, -0.2776532711191075, -0.6615321044132733, -0.6615321044132735,
    0.0000000000000002, 2.4336772933029756, -0.5553065422382126, -0.5553065422382162, -0.2776532711191089,
    0.0000000000000003, -0.2776532711191083, -0.5553065422382126, 2.4336772933029738, -1.3230642088265447,
    0.0000000000000013, -0.2776532711191075, -0.2776532711191076, -0.5553065422382162, -1.3230642088265447,
    2.4336772933029751]


def test_laplacian_order2_local_matrix(golden_meshes):
    # fem_operators_test.cpp:41-100: L = -laplacian<FEM>() on c_shaped cell 175, Integrator<FEM,2,2> (6-point rule)
    pts, els, _ = golden_meshes("c_shaped")
    A = orc.local_matrix(2, 2, pts[els[175]], [(orc.LAPLACIAN, -1.0)])
    got = A.ravel()
    for a, b in zip(got, GOLDEN_P2_STIFF):
        assert almost_equal(a, b)  # the reference's own criterion
    assert np.max(np.abs(got - np.array(GOLDEN_P2_STIFF))) < 5e-15  # and far tighter than that


# ---- geometry / quadrature ---------------------------------------------------------------------------------------

def test_simplex_measures():
    # simplex_test.cpp:27-34, 89-97, 57-63
    assert almost_equal(orc.cell_geometry([[0, 0], [0.5, 0], [0, 0.8]])[2], 0.5 * 0.8 / 2)
    assert almost_equal(orc.cell_geometry([[0, 0, 0], [0.4, 0.2, 0], [0, 0.8, 0.6], [0.4, 0.6, 0.8]])[2],
                        0.0266666666666666)
    assert almost_equal(orc.cell_geometry([[0, 0, 0], [0.5, 0.2, 0], [0, 0.8, 0.6]])[2], 0.25709920264364883)


def test_inverse_jacobian():
    rng = np.random.default_rng(0)
    for M in (2, 3):
        v = rng.random((M + 1, M))
        J, invJ, meas = orc.cell_geometry(v)
        assert np.allclose(invJ @ J, np.eye(M), atol=1e-12)
        assert np.isclose(meas, abs(np.linalg.det(J)) / (2 if M == 2 else 6))


@pytest.mark.parametrize("M,Ks", [(2, [1, 3, 6, 7, 12]), (3, [1, 4, 5, 11])])
def test_quadrature_tables_consistent(M, Ks):
    # integration_test.cpp:112-125: every table of a dimension integrates the same polynomial alike
    vals = []
    for K in Ks:
        n, w = orc.quadrature_table(M, K)
        assert abs(w.sum() - 1.0) < 1e-12
        vals.append((w * (1.0 + n[:, 0] + 2 * n[:, -1])).sum())
    assert np.ptp(vals) < 1e-12


def test_integrate_one_unit_square(golden_meshes):
    # integration_test.cpp:72-80: the measure of unit_square is 1
    pts, els, _ = golden_meshes("unit_square")
    assert almost_equal(orc.integrate_one(1, pts, els), 1.0)
    assert almost_equal(orc.integrate_one(2, pts, els), 1.0)


def test_integrate_linear_field_cell(golden_meshes):
    # integration_test.cpp:46-70 restated: int_e (x + y) = measure * mean of vertex values
    pts, els, _ = golden_meshes("unit_square")
    e = 1234
    q = orc.quadrature_nodes(1, pts, els)
    _, w = orc.quadrature(2, 1)
    meas = orc.cell_geometry(pts[els[e]])[2]
    val = ((q[3 * e:3 * e + 3, 0] + q[3 * e:3 * e + 3, 1]) * w).sum() * meas
    assert almost_equal(val, meas * pts[els[e]].sum(axis=0).sum() / 3)


# ---- topology / dofs ---------------------------------------------------------------------------------------------

def test_unit_square_topology(golden_meshes):
    # mesh_loader.h:35: 3600 points, 6962 elements, 10561 edges; SURVEY Appendix A.9: 14161 P2 dofs, 472 boundary
    pts, els, bnd = golden_meshes("unit_square")
    ce, edges, eb = orc.enumerate_edges(els)
    assert edges.shape[0] == 10561
    dofs, n_dofs, bd = orc.enumerate_dofs(2, pts.shape[0], els, bnd)
    assert n_dofs == 14161 and int(bd.sum()) == 472 and int(bnd.sum()) == 236
    # first-occurrence order: the first cell holds edges 0,1,2 in pair order (0,1),(0,2),(1,2)
    assert ce[0].tolist() == [0, 1, 2]
    assert dofs[0].tolist() == els[0].tolist() + [3600, 3601, 3602]


def test_edge_ids_match_hash_scan(golden_meshes):
    # literal python restatement of the unordered_map scan (triangulation.h:167-192) vs the sort-based oracle
    pts, els, _ = golden_meshes("c_shaped")
    seen, ids = {}, []
    for c in els:
        row = []
        for a, b in ((0, 1), (0, 2), (1, 2)):
            k = tuple(sorted((int(c[a]), int(c[b]))))
            if k not in seen:
                seen[k] = len(seen)
            row.append(seen[k])
        ids.append(row)
    ce, edges, eb = orc.enumerate_edges(els)
    assert ce.tolist() == ids
    assert [tuple(e) for e in edges.tolist()] == list(seen.keys())


def test_edge_ids_3d_match_hash_scan(golden_meshes):
    # literal restatement of triangulation.h:347-377 (edges numbered inside each NEW face, sorted triple)
    pts, els, bnd = golden_meshes("unit_sphere")
    faces, edges = set(), {}
    for c in els:
        for f in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)):
            face = tuple(sorted(int(c[k]) for k in f))
            if face in faces:
                faces.discard(face)
                continue
            faces.add(face)
            for a, b in ((0, 1), (0, 2), (1, 2)):
                k = (face[a], face[b])
                if k not in edges:
                    edges[k] = len(edges)
    ce, e_or, eb = orc.enumerate_edges(els, bnd)
    assert [tuple(e) for e in e_or.tolist()] == list(edges.keys())
    for c, row in zip(els[:200], ce[:200]):
        for (a, b), eid in zip(((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)), row):
            assert edges[tuple(sorted((int(c[a]), int(c[b]))))] == eid


# ---- global assembly semantics -----------------------------------------------------------------------------------

def test_structural_nnz_keeps_explicit_zeros(golden_meshes):
    # P1 nnz = n_nodes + 2 n_edges = 24722 on unit_square (nothing pruned)
    pts, els, _ = golden_meshes("unit_square")
    outer, inner, val = orc.assemble_operator(1, pts, els, els, pts.shape[0], [(orc.LAPLACIAN, -1.0)], True)
    assert inner.size == 3600 + 2 * 10561
    assert np.all(np.diff(outer) > 0)
    for j in (0, 17, 3599):
        assert np.all(np.diff(inner[outer[j]:outer[j + 1]]) > 0)  # sorted inner indices
    A = csc(outer, inner, val, 3600)
    assert abs(A - A.T).max() == 0.0  # selfadjointView<Lower> mirrors bit-exactly
    assert np.abs(A @ np.ones(3600)).max() < 1e-11  # stiffness annihilates constants


def test_symmetric_and_full_paths_agree(golden_meshes):
    pts, els, _ = golden_meshes("unit_square_16")
    n = pts.shape[0]
    t = [(orc.LAPLACIAN, -1.0), (orc.REACTION, 1.0, [2.5])]
    o1, i1, v1 = orc.assemble_operator(1, pts, els, els, n, t, True)
    o2, i2, v2 = orc.assemble_operator(1, pts, els, els, n, t, False)
    assert np.array_equal(o1, o2) and np.array_equal(i1, i2)
    assert np.max(np.abs(v1 - v2)) < 1e-13


def test_mass_matrix_sums_to_area(golden_meshes):
    pts, els, bnd = golden_meshes("unit_square")
    for R in (1, 2):
        dofs, n_dofs, _ = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
        o, i, v = orc.assemble_operator(R, pts, els, dofs, n_dofs, [(orc.REACTION, 1.0, [1.0])], True)
        assert almost_equal(v.sum(), 1.0)


# ---- fem_pde_test.cpp end-to-end cases (SparseLU -> SuperLU/COLAMD) ----------------------------------------------

def solve_pde(R, pts, els, bnd, terms, symmetric, f_fn, g_fn):
    dofs, n_dofs, bd = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
    xy = orc.dofs_coords(R, pts, els, dofs, n_dofs)
    o, i, v = orc.assemble_operator(R, pts, els, dofs, n_dofs, terms, symmetric)
    q = orc.quadrature_nodes(R, pts, els)
    b = orc.assemble_forcing(R, pts, els, dofs, n_dofs, f_fn(q))
    mo, mi, mv = orc.assemble_operator(R, pts, els, dofs, n_dofs, [(orc.REACTION, 1.0, [1.0])], True)
    g = g_fn(xy)
    orc.set_dirichlet(o, i, v, bd, g, b)
    lu = spla.splu(csc(o, i, v, n_dofs), permc_spec="COLAMD")
    u = lu.solve(b)
    return u, xy, csc(mo, mi, mv, n_dofs)


def l2_error(mass, u_ex, u):  # fem_pde_test.cpp:72-74
    err = u_ex - u
    return float((mass @ (err * err)).sum())


def test_pde_laplacian_order1(golden_meshes):
    # fem_pde_test.cpp:43-75
    pts, els, bnd = golden_meshes("unit_square")
    ex = lambda x: x[:, 0] + x[:, 1]
    u, xy, mass = solve_pde(1, pts, els, bnd, [(orc.LAPLACIAN, -1.0)], True, lambda q: np.zeros(q.shape[0]), ex)
    assert l2_error(mass, ex(xy), u) < DOUBLE_TOLERANCE


def test_pde_laplacian_order2_callable_force(golden_meshes):
    # fem_pde_test.cpp:78-107: u = 1 - x^2 - y^2, f = 4
    pts, els, bnd = golden_meshes("unit_square")
    ex = lambda x: 1.0 - x[:, 0] ** 2 - x[:, 1] ** 2
    u, xy, mass = solve_pde(2, pts, els, bnd, [(orc.LAPLACIAN, -1.0)], True, lambda q: np.full(q.shape[0], 4.0), ex)
    assert l2_error(mass, ex(xy), u) < DOUBLE_TOLERANCE


def _advdiff_exact():
    pi, alpha, gamma = np.pi, 1.0, np.pi
    l1 = -alpha / 2 - np.sqrt((alpha / 2) ** 2 + pi * pi)
    l2 = -alpha / 2 + np.sqrt((alpha / 2) ** 2 + pi * pi)
    p = (1 - np.exp(l2)) / (np.exp(l1) - np.exp(l2))
    ex = lambda x: -gamma / (pi * pi) * (p * np.exp(l1 * x[:, 0]) + (1 - p) * np.exp(l2 * x[:, 0]) - 1.0) * np.sin(
        pi * x[:, 1])
    f = lambda q: gamma * np.sin(pi * q[:, 1])
    return ex, f


@pytest.mark.parametrize("R,tol", [(1, 1e-5), (2, 1e-7)])
def test_pde_advection_diffusion(golden_meshes, R, tol):
    # fem_pde_test.cpp:113-166 (order 1, < 1e-5) and :172-212 (order 2, < 1e-7)
    pts, els, bnd = golden_meshes("unit_square")
    ex, f = _advdiff_exact()
    terms = [(orc.LAPLACIAN, -1.0), (orc.ADVECTION, 1.0, [-1.0, 0.0])]
    u, xy, mass = solve_pde(R, pts, els, bnd, terms, False, f, lambda x: np.zeros(x.shape[0]))
    assert l2_error(mass, ex(xy), u) < tol


def test_advection_orientation_is_pinned(golden_meshes):
    # SURVEY section 7 hard part 1: the transposed orientation (row = trial) fails fem_pde_test.cpp:165
    pts, els, bnd = golden_meshes("unit_square")
    ex, f = _advdiff_exact()
    n = pts.shape[0]
    o, i, v = orc.assemble_operator(1, pts, els, els, n, [(orc.LAPLACIAN, -1.0), (orc.ADVECTION, 1.0, [-1.0, 0.0])],
                                    False)
    At = csc(o, i, v, n).T.tocsc()
    At.sort_indices()
    b = orc.assemble_forcing(1, pts, els, els, n, f(orc.quadrature_nodes(1, pts, els)))
    vt = At.data.copy()
    orc.set_dirichlet(At.indptr, At.indices, vt, bnd, np.zeros(n), b)
    u = spla.splu(csc(At.indptr, At.indices, vt, n), permc_spec="COLAMD").solve(b)
    mo, mi, mv = orc.assemble_operator(1, pts, els, els, n, [(orc.REACTION, 1.0, [1.0])], True)
    assert l2_error(csc(mo, mi, mv, n), ex(pts), u) > 1e-5


# ---- Krylov oracles vs LU ----------------------------------------------------------------------------------------

def test_cg_matches_lu_with_dirichlet_rows(golden_meshes):
    # CG on the row-replaced matrix is valid iff x0 carries the boundary values
    pts, els, bnd = golden_meshes("unit_square")
    n = pts.shape[0]
    o, i, v = orc.assemble_operator(1, pts, els, els, n, [(orc.LAPLACIAN, -1.0)], True)
    b = orc.assemble_forcing(1, pts, els, els, n, np.ones(els.shape[0] * 3))
    g = pts[:, 0] + pts[:, 1]
    orc.set_dirichlet(o, i, v, bnd, g, b)
    A = csc(o, i, v, n)
    u_lu = spla.splu(A, permc_spec="COLAMD").solve(b)
    Ar = A.tocsr()
    Ar.sort_indices()
    x0 = np.where(bnd > 0, g, 0.0)
    u, it, rel = orc.cg(Ar.indptr, Ar.indices, Ar.data, b, x0, rtol=1e-12)
    assert rel <= 1e-12 and it < 1000
    assert np.linalg.norm(u - u_lu) / np.linalg.norm(u_lu) < 1e-8


def test_bicgstab_matches_lu(golden_meshes):
    pts, els, bnd = golden_meshes("unit_square_32")
    n = pts.shape[0]
    ex, f = _advdiff_exact()
    o, i, v = orc.assemble_operator(1, pts, els, els, n, [(orc.LAPLACIAN, -1.0), (orc.ADVECTION, 1.0, [-1.0, 0.0])],
                                    False)
    b = orc.assemble_forcing(1, pts, els, els, n, f(orc.quadrature_nodes(1, pts, els)))
    orc.set_dirichlet(o, i, v, bnd, np.zeros(n), b)
    A = csc(o, i, v, n)
    u_lu = spla.splu(A, permc_spec="COLAMD").solve(b)
    Ar = A.tocsr()
    Ar.sort_indices()
    u, it, rel = orc.bicgstab(Ar.indptr, Ar.indices, Ar.data, b, np.zeros(n), rtol=1e-12)
    assert np.linalg.norm(u - u_lu) / np.linalg.norm(u_lu) < 1e-8


# ---- next-row N1: Psi evaluation pinned by the reference's .mtx fixtures (lagrangian_basis_test.cpp:200-238) ------------
def _psi_fixtures():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "psi.npz"))


def _dense(shape, r, c, v):
    import scipy.sparse as sp
    return sp.coo_matrix((v, (r, c)), shape=tuple(int(x) for x in shape)).toarray()   # duplicates are summed


def _almost_equal(a, b, eps=1e-7):
    """test/src/utils/utils.h:44-48 (DOUBLE_TOLERANCE = 1e-7, constants.h:11) -- the reference's own acceptance test"""
    d = np.max(np.abs(a - b))
    return d < eps or d < max(np.max(np.abs(a)), np.max(np.abs(b))) * eps


@pytest.mark.parametrize("R", [1, 2])
def test_pointwise_evaluation_matches_reference_mtx(golden_meshes, R):
    z = _psi_fixtures()
    pts, els, bnd = golden_meshes("c_shaped")
    dofs, n_dofs, _ = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
    locs = z["c_shaped/locs"]
    ids, cols, vals = orc.eval_pointwise(R, pts, els, dofs, locs)
    assert (ids >= 0).all()
    name = f"lagrangian_pointwise_eval_order{R}"
    assert tuple(z[name + "/shape"]) == (locs.shape[0], n_dofs)
    rows = np.repeat(np.arange(locs.shape[0]), cols.shape[1])
    got = _dense((locs.shape[0], n_dofs), rows, cols.ravel(), vals.ravel())
    want = _dense(z[name + "/shape"], z[name + "/rows"], z[name + "/cols"], z[name + "/vals"])
    assert _almost_equal(got, want)
    assert np.max(np.abs(got - want)) < 1e-13          # far inside the reference's 1e-7
    assert np.allclose(got.sum(axis=1), 1.0, atol=1e-13)   # partition of unity


@pytest.mark.parametrize("R", [1, 2])
def test_areal_evaluation_matches_reference_mtx(golden_meshes, R):
    z = _psi_fixtures()
    pts, els, bnd = golden_meshes("quasi_circle")
    dofs, n_dofs, _ = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
    inc = z["quasi_circle/incidence"]
    rows, cols, vals, D = orc.eval_areal(R, pts, els, dofs, inc)
    name = f"lagrangian_areal_eval_order{R}"
    assert tuple(z[name + "/shape"]) == (inc.shape[0], n_dofs)
    got = _dense((inc.shape[0], n_dofs), rows, cols, vals)
    want = _dense(z[name + "/shape"], z[name + "/rows"], z[name + "/cols"], z[name + "/vals"])
    assert _almost_equal(got, want)
    assert np.max(np.abs(got - want)) < 1e-9
    assert np.allclose(got.sum(axis=1), 1.0, atol=1e-12)   # each row averages a partition of unity
    assert np.all(D > 0)


def test_locate_outside_and_shared_points(golden_meshes):
    pts, els, bnd = golden_meshes("unit_square")
    far = np.array([[2.0, 2.0], [-0.5, 0.5]])
    assert np.array_equal(orc.locate(pts, els, far), [-1, -1])
    # a mesh node is shared by several cells: the smallest cell id wins
    ids = orc.locate(pts, els, pts[:50])
    for i, e in enumerate(ids):
        assert e == np.nonzero((els == i).any(axis=1))[0].min()


# ---- next-row N2: the parabolic driver, pinned by fem_pde_test.cpp:222-285 ---------------------------------------------
def test_parabolic_isotropic_order2_threshold(golden_meshes):
    """dt(u) - lap u = f on unit_square, P2, 101 time steps, u = sin(2 pi x) sin(2 pi y) exp(-t):
    max_j (mass * err_j^2).sum() < 1e-7 (the reference's own acceptance test)."""
    from parabolic_ref import parabolic_reference
    pts, els, bnd = golden_meshes("unit_square")
    pi = np.pi
    times = np.linspace(0.0, 1.0, 101)
    u_fn = lambda x, t: np.sin(2 * pi * x[:, 0]) * np.sin(2 * pi * x[:, 1]) * np.exp(-t)
    f_fn = lambda x, t: (8 * pi * pi - 1.0) * np.sin(2 * pi * x[:, 0]) * np.sin(2 * pi * x[:, 1]) * np.exp(-t)
    sol, xy, q, mass = parabolic_reference(2, pts, els, bnd, times, u_fn, f_fn)
    errs = [float((mass @ ((u_fn(xy, t) - sol[:, j]) ** 2)).sum()) for j, t in enumerate(times)]
    assert max(errs) < 1e-7


def test_all_core_variant_is_bit_identical(golden_meshes):
    """bench.py --impl reference uses the OpenMP build of the oracle's assembly: same triplet list, same matrix."""
    pts, els, bnd = golden_meshes("unit_sphere")
    for R in (1, 2):
        dofs, n_dofs, _ = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
        for terms, sym in (([(orc.LAPLACIAN, -1.0)], True),
                           ([(orc.LAPLACIAN, -1.0), (orc.ADVECTION, 1.0, [1.0, 0.0, -1.0]), (orc.REACTION, 2.0, [1.0])], False)):
            a = orc.assemble_operator(R, pts, els, dofs, n_dofs, terms, sym)
            for nt in (1, 3, 8):
                b = orc.assemble_operator_mt(R, pts, els, dofs, n_dofs, terms, sym, n_threads=nt)
                assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_parabolic_isotropic_order1_convergence(fdb):
    """fem_pde_test.cpp:295-368: P1, 31 time steps, meshes unit_square_{16,32,64,128} (the synthetic generator reproduces
    those files node for node, checked against the reference's files when the fixtures were made): the L2 error at the
    final time must fall with order 2, floor(log2(e_k / e_{k+1})) == 2 for every refinement."""
    from parabolic_ref import parabolic_reference
    pi = np.pi
    times = np.linspace(0.0, 1.0, 31)
    u_fn = lambda x, t: np.sin(2 * pi * x[:, 0]) * np.sin(2 * pi * x[:, 1]) * np.exp(-t)
    f_fn = lambda x, t: (8 * pi * pi - 1.0) * np.sin(2 * pi * x[:, 0]) * np.sin(2 * pi * x[:, 1]) * np.exp(-t)
    errs = []
    for N in (16, 32, 64, 128):
        pts, els, bnd = fdb.meshes.unit_square(N)
        sol, xy, q, mass = parabolic_reference(1, pts, els, bnd, times, u_fn, f_fn)
        e = u_fn(xy, times[-1]) - sol[:, -1]
        errs.append(np.sqrt(float((mass @ (e * e)).sum())))
    orders = [np.log2(errs[k] / errs[k + 1]) for k in range(3)]
    assert all(np.floor(o) == 2 for o in orders), (errs, orders)


@pytest.mark.parametrize("mesh", ["unit_square", "unit_sphere", "c_shaped", "surface"])
def test_point_location_sampled_like_the_reference(golden_meshes, mesh):
    """point_location_test.cpp:38-50 with MeshLoader::sample (mesh_loader.h:88-121): 100 random cells, one random point
    inside each (convex combination of the vertices); locate must return exactly that cell."""
    pts, els, _ = golden_meshes(mesh)
    rng = np.random.default_rng(123)
    ids = rng.integers(0, els.shape[0], 100)
    M = els.shape[1] - 1
    v = pts[els[ids]]
    t = rng.random(100)[:, None]
    p = t * v[:, 0] + (1 - t) * v[:, 1]
    for j in range(1, M):
        t = rng.random(100)[:, None]
        p = (1 - t) * v[:, 1 + j] + t * p
    assert np.array_equal(orc.locate(pts, els, p), ids)


def test_pointwise_evaluation_on_the_surface_mesh(golden_meshes):
    """Triangulation<2,3>: location needs the supporting-plane test (simplex.h:116-118); Psi rows at points of the surface
    are a partition of unity and reproduce the embedding coordinates (P1 and P2 both contain the affine functions)."""
    pts, els, bnd = golden_meshes("surface")
    rng = np.random.default_rng(5)
    ids = rng.integers(0, els.shape[0], 200)
    w = rng.dirichlet(np.ones(3), 200)
    p = np.einsum("ik,ikd->id", w, pts[els[ids]])
    for R in (1, 2):
        dofs, n_dofs, _ = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
        got, cols, vals = orc.eval_pointwise(R, pts, els, dofs, p)
        assert np.array_equal(got, ids)
        assert np.allclose(vals.sum(axis=1), 1.0, atol=1e-13)
        xc = orc.dofs_coords(R, pts, els, dofs, n_dofs)
        assert np.max(np.abs(np.einsum("ih,ihd->id", vals, xc[cols]) - p)) < 1e-13
    off = p + np.array([0.0, 0.0, 1e-9])          # a point off the surface is in no cell
    assert (orc.locate(pts, els, off) == -1).all()


# ---- next-row N3: mesh topology (Triangulation constructors, triangulation.h:143-196, 319-399) ------------------------
@pytest.mark.parametrize("mesh", ["c_shaped", "unit_square", "surface", "quasi_circle", "unit_sphere"])
def test_topology_oracle_matches_reference_fixtures(golden_meshes, mesh):
    """neigh.csv pins the neighbour table exactly (column j = cell across the facet opposite to vertex j, -1 = boundary);
    edges.csv pins the facet SET (its row order is the mesh generator's, not the constructor's first-occurrence order)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    pts, els, bnd = golden_meshes(mesh)
    t = orc.mesh_topology(els, bnd)
    assert np.array_equal(t["neighbors"], z[mesh + "/neigh"])
    file_facets = np.sort(z[mesh + "/facets_file"][:, :els.shape[1] - 1], axis=1)
    assert t["n_facets"] == file_facets.shape[0] == int(z[mesh + "/n_edges_file"])
    assert set(map(tuple, t["facets"])) == set(map(tuple, file_facets))
    # a facet is on the boundary iff one cell holds it; its nodes are then boundary nodes of the fixture
    assert np.array_equal(t["facet_boundary"] == 1, t["facet_to_cells"][:, 1] == -1)
    assert np.all(np.asarray(bnd).ravel()[t["facets"][t["facet_boundary"] == 1]] == 1)


@pytest.mark.parametrize("mesh", ["c_shaped", "unit_sphere"])
def test_topology_oracle_equals_literal_python_scan(golden_meshes, mesh):
    """the C restatement against a dict-based transcription of the constructor loops (ids in first-occurrence order)"""
    pts, els, bnd = golden_meshes(mesh)
    M = els.shape[1] - 1
    pat = [(0, 1), (0, 2), (1, 2)] if M == 2 else [(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)]
    fmap, facets, f2c, c2f, emap, edges, f2e = {}, [], [], np.zeros_like(els), {}, [], []
    e2c = {}
    for i, c in enumerate(els):
        for j, p in enumerate(pat):
            f = tuple(sorted(int(c[k]) for k in p))
            if f not in fmap:
                fmap[f] = (len(facets), i)
                c2f[i, j] = len(facets)
                facets.append(f)
                f2c.append([i, -1])
                if M == 3:
                    row = []
                    for a, b in [(0, 1), (0, 2), (1, 2)]:
                        e = (f[a], f[b])
                        if e not in emap:
                            emap[e] = len(edges)
                            edges.append(e)
                        row.append(emap[e])
                        e2c.setdefault(emap[e], set()).add(i)
                    f2e.append(row)
            else:
                h, k = fmap.pop(f)
                c2f[i, j] = h
                f2c[h][1] = i
                if M == 3:
                    for e in f2e[h]:
                        e2c[e].add(i)
    t = orc.mesh_topology(els, bnd)
    assert np.array_equal(t["facets"], np.array(facets)) and np.array_equal(t["cell_to_facets"], c2f)
    assert np.array_equal(t["facet_to_cells"], np.array(f2c))
    if M == 3:
        assert np.array_equal(t["edges"], np.array(edges)) and np.array_equal(t["face_to_edges"], np.array(f2e))
        for e in range(len(edges)):
            assert sorted(e2c[e]) == list(t["edge_cells"][t["edge_cell_ptr"][e]:t["edge_cell_ptr"][e + 1]])
        bn = np.asarray(bnd).ravel()
        assert np.array_equal(t["edge_boundary"], (bn[t["edges"][:, 0]] & bn[t["edges"][:, 1]]).astype(np.uint8))
