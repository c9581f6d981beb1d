"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): sparsity pattern bit-exact, matrix entries within 1e-12 relative (absolute floor
1e-12 * |diagonal| for structural zeros), solutions within 1e-8 relative L2 of the direct (SuperLU/COLAMD) solve.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import entry_tolerance
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

ENTRY_RTOL = 1e-12
SOLUTION_RTOL = 1e-8


def orc_terms(expr, N):
    out = []
    for k, s, c, sv in expr.leaves:
        if c is None:
            out.append((k, s))
        elif k == orc.DIFFUSION and not sv:
            out.append((k, s, np.asarray(c).reshape(N, N, order="F"), 0))  # Terms() re-flattens column-major
        else:
            out.append((k, s, c, int(sv)))
    return out


def check_operator(fdb, nodes, cells, R, dofs, n_dofs, expr, symmetric=None):
    mesh = fdb.Triangulation(nodes, cells, np.zeros(nodes.shape[0], np.uint8))
    asm = fdb.Assembler(mesh, R, n_dofs, dofs)
    sym = expr.is_symmetric if symmetric is None else symmetric
    outer, inner, val = asm.discretize_operator(expr, sym)
    o, i, v = orc.assemble_operator(R, nodes, cells, dofs, n_dofs, orc_terms(expr, nodes.shape[1]), sym)
    assert np.array_equal(outer, o), "outer index array differs"
    assert np.array_equal(inner, i), "inner index array differs"
    tol = entry_tolerance(o, i, v, ENTRY_RTOL)
    bad = np.abs(val - v) > tol
    assert not bad.any(), f"{bad.sum()} entries differ, worst {np.max(np.abs(val - v) / np.maximum(tol, 1e-300)):.3g} x tol"
    return asm, (outer, inner, val)


# ---- A8: operator assembly ---------------------------------------------------------------------------------------

def test_p1_laplacian_unit_square(fdb, golden_meshes):
    pts, els, _ = golden_meshes("unit_square")
    _, (outer, inner, val) = check_operator(fdb, pts, els, 1, els, pts.shape[0], -fdb.laplacian())
    assert inner.size == 3600 + 2 * 10561  # explicit zeros kept


@pytest.mark.parametrize("mesh", ["c_shaped", "unit_square"])
def test_p1_operators_2d(fdb, golden_meshes, mesh):
    pts, els, _ = golden_meshes(mesh)
    n = pts.shape[0]
    check_operator(fdb, pts, els, 1, els, n, fdb.reaction(1.0))
    check_operator(fdb, pts, els, 1, els, n, -fdb.laplacian() + fdb.advection([-1.0, 0.0]))
    check_operator(fdb, pts, els, 1, els, n, -fdb.diffusion([[2.0, 0.3], [0.3, 1.0]]) + 0.5 * fdb.reaction(3.0))
    check_operator(fdb, pts, els, 1, els, n, -fdb.laplacian() + fdb.reaction(2.0), symmetric=False)


@pytest.mark.parametrize("mesh", ["c_shaped", "unit_square"])
def test_p2_operators_2d(fdb, golden_meshes, mesh):
    pts, els, bnd = golden_meshes(mesh)
    dofs, n_dofs, _ = orc.enumerate_dofs(2, pts.shape[0], els, bnd)
    check_operator(fdb, pts, els, 2, dofs, n_dofs, -fdb.laplacian())
    check_operator(fdb, pts, els, 2, dofs, n_dofs, fdb.reaction(1.0))
    check_operator(fdb, pts, els, 2, dofs, n_dofs,
                   -fdb.laplacian() + fdb.advection([-1.0, 0.5]) + fdb.reaction(1.0))


# ---- next-row N3: manifold cells, Triangulation<2,3> (simplex.h:189-193) ------------------------------------------------
@pytest.mark.parametrize("R", [1, 2])
def test_surface_operators(fdb, golden_meshes, R):
    pts, els, bnd = golden_meshes("surface")
    assert pts.shape[1] == 3 and els.shape[1] == 3
    dofs, n_dofs, _ = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
    K = [[2.0, 0.3, 0.1], [0.3, 1.0, -0.2], [0.1, -0.2, 1.5]]
    _, (o, i, v) = check_operator(fdb, pts, els, R, dofs, n_dofs, -fdb.laplacian())       # Laplace-Beltrami stiffness
    A = sp.csc_matrix((v, i, o), shape=(n_dofs, n_dofs))
    assert np.max(np.abs(A.sum(axis=1))) < 1e-12                                          # constants are in the kernel
    _, (o, i, v) = check_operator(fdb, pts, els, R, dofs, n_dofs, fdb.reaction(1.0))
    area = sum(orc.cell_geometry(pts[c])[2] for c in els)
    assert abs(v.sum() - area) < 1e-12 * area                                             # sum of the mass matrix = area
    check_operator(fdb, pts, els, R, dofs, n_dofs, -fdb.diffusion(K) + 0.5 * fdb.reaction(3.0))
    check_operator(fdb, pts, els, R, dofs, n_dofs, -fdb.laplacian() + fdb.advection([1.0, -0.5, 0.25]) + fdb.reaction(1.0))
    check_operator(fdb, pts, els, R, dofs, n_dofs, -fdb.laplacian() + fdb.reaction(2.0), symmetric=False)


def test_surface_reaction_diffusion_solve(fdb, golden_meshes):
    """-Laplace-Beltrami u + u = f on the surface mesh (closed-form-free check: GPU CG against SuperLU on the oracle's
    matrix and load vector, Dirichlet data on the boundary nodes)."""
    pts, els, bnd = golden_meshes("surface")
    n = pts.shape[0]
    mesh = fdb.Triangulation(pts, els, bnd)
    expr = -fdb.laplacian() + fdb.reaction(1.0)
    pde = fdb.PDE(mesh, expr, 1, forcing=lambda q: np.cos(2 * q[:, 0]) + q[:, 2], solver=fdb.SolverOptions("cg", rtol=1e-12))
    g = pts[:, 0] - pts[:, 1] * pts[:, 2]
    pde.set_dirichlet_bc(g)
    pde.init()
    pde.solve()
    assert pde.success
    q = orc.quadrature_nodes(1, pts, els)
    o, i, v = orc.assemble_operator(1, pts, els, els, n, orc_terms(expr, 3), True)
    b = orc.assemble_forcing(1, pts, els, els, n, np.cos(2 * q[:, 0]) + q[:, 2])
    orc.set_dirichlet(o, i, v, bnd, g, b)
    u = spla.splu(sp.csc_matrix((v, i, o), shape=(n, n)), permc_spec="COLAMD").solve(b)
    assert np.linalg.norm(pde.solution() - u) <= SOLUTION_RTOL * np.linalg.norm(u)


def test_golden_p2_local_stiffness_through_the_gpu(fdb, golden_meshes):
    # fem_operators_test.cpp:41-100 via a one-cell mesh: the global matrix IS the local matrix
    from test_oracle_golden import GOLDEN_P2_STIFF
    pts, els, _ = golden_meshes("c_shaped")
    nodes = pts[els[175]]
    cells = np.array([[0, 1, 2]], dtype=np.int32)
    dofs = np.arange(6, dtype=np.int32).reshape(1, 6)
    mesh = fdb.Triangulation(nodes, cells, np.zeros(3, np.uint8))
    outer, inner, val = fdb.Assembler(mesh, 2, 6, dofs).discretize_operator(-fdb.laplacian(), symmetric=False)
    A = sp.csc_matrix((val, inner, outer), shape=(6, 6)).toarray()
    assert np.max(np.abs(A.ravel() - np.array(GOLDEN_P2_STIFF))) < 1e-13


def test_p1_3d_unit_sphere_mixed_orientation(fdb, golden_meshes):
    # 1395 of 2775 tets have det J < 0: |det| must be used
    pts, els, _ = golden_meshes("unit_sphere")
    n = pts.shape[0]
    check_operator(fdb, pts, els, 1, els, n, -fdb.laplacian())
    check_operator(fdb, pts, els, 1, els, n, fdb.reaction(1.0))
    K = np.array([[1.0, 0.2, 0.0], [0.2, 2.0, 0.1], [0.0, 0.1, 3.0]])
    check_operator(fdb, pts, els, 1, els, n, -fdb.diffusion(K) + fdb.advection([0.3, -1.0, 0.5]) + fdb.reaction(0.7))


@pytest.mark.parametrize("n", [4, 8, 16, 32])     # SURVEY 8(d) ladder n in {8, 16, 32}
def test_p1_3d_kuhn_cube_ladder(fdb, n):
    nodes, cells, bnd = fdb.meshes.unit_cube(n)
    check_operator(fdb, nodes, cells, 1, cells, nodes.shape[0], -fdb.laplacian())
    jn = fdb.meshes.jitter(nodes, bnd, 1.0 / n)
    check_operator(fdb, jn, cells, 1, cells, nodes.shape[0], -fdb.laplacian() + fdb.reaction(1.0))


@pytest.mark.parametrize("N", [16, 64, 256])      # SURVEY 8(d) ladder N in {16, 64, 256}
def test_p1_2d_square_ladder(fdb, N):
    nodes, cells, bnd = fdb.meshes.unit_square(N)
    check_operator(fdb, nodes, cells, 1, cells, nodes.shape[0], -fdb.laplacian())
    check_operator(fdb, nodes, cells, 1, cells, nodes.shape[0], fdb.reaction(1.0))


def test_p2_3d_extension(fdb, golden_meshes):
    # A10: not in the reference; parity is against the oracle's statement of the same convention
    pts, els, bnd = golden_meshes("unit_sphere")
    dofs, n_dofs, _ = orc.enumerate_dofs(2, pts.shape[0], els, bnd)
    check_operator(fdb, pts, els, 2, dofs, n_dofs, -fdb.laplacian() + fdb.reaction(1.0))
    check_operator(fdb, pts, els, 2, dofs, n_dofs, -fdb.laplacian() + fdb.advection([1.0, 0.0, -1.0]))
    nodes, cells, bnd = fdb.meshes.unit_cube(4)
    dofs, n_dofs, _ = orc.enumerate_dofs(2, nodes.shape[0], cells, bnd)
    _, (o, i, v) = check_operator(fdb, nodes, cells, 2, dofs, n_dofs, fdb.reaction(1.0))
    assert abs(v.sum() - 1.0) < 1e-12  # sum of the mass matrix = volume


def test_space_varying_coefficients(fdb, golden_meshes):
    # Discretized{Matrix,Vector,Scalar}Field rows nq*e+q (integrator.h:98-101)
    pts, els, bnd = golden_meshes("c_shaped")
    rng = np.random.default_rng(1)
    for R, nq in ((1, 3), (2, 6)):
        dofs, n_dofs, _ = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
        rows = els.shape[0] * nq
        K = rng.random((rows, 4)) + np.array([2.0, 0.0, 0.0, 2.0])
        K[:, 1] = K[:, 2]
        b = rng.standard_normal((rows, 2))
        c = rng.random(rows)
        check_operator(fdb, pts, els, R, dofs, n_dofs, -fdb.diffusion(K) + fdb.reaction(c))
        check_operator(fdb, pts, els, R, dofs, n_dofs, -fdb.laplacian() + fdb.advection(b))


def test_assembly_is_bit_reproducible(fdb):
    nodes, cells, bnd = fdb.meshes.unit_cube(12)
    jn = fdb.meshes.jitter(nodes, bnd, 1.0 / 12)
    mesh = fdb.Triangulation(jn, cells, bnd)
    asm = fdb.Assembler(mesh, 1, nodes.shape[0], cells)
    a = asm.discretize_operator(-fdb.laplacian())[2]
    for _ in range(3):
        assert a.tobytes() == asm.discretize_operator(-fdb.laplacian())[2].tobytes()
    other = fdb.Assembler(mesh, 1, nodes.shape[0], cells).discretize_operator(-fdb.laplacian())[2]
    assert a.tobytes() == other.tobytes()


@pytest.mark.parametrize("case", ["p1_3d", "p1_2d", "p2_2d", "p2_3d", "surface_p2"])
def test_rowwise_pattern_build_equals_sort_build(fdb, golden_meshes, case, monkeypatch):
    # the row-wise pattern build (default) and the sort of all emitted triplets (FDB_PATTERN_SORT=1) must produce the same
    # pattern, the same scatter map and the same segment order: values bit-identical on both assembly paths
    if case == "p1_3d":
        nodes, cells, bnd = fdb.meshes.unit_cube(9)
        nodes = fdb.meshes.jitter(nodes, bnd, 1.0 / 9)
        R = 1
    elif case == "p2_3d":
        nodes, cells, bnd = fdb.meshes.unit_cube(5)
        nodes = fdb.meshes.jitter(nodes, bnd, 1.0 / 5)
        R = 2
    elif case == "surface_p2":
        nodes, cells, bnd = golden_meshes("surface")
        R = 2
    else:
        nodes, cells, bnd = golden_meshes("unit_square")
        R = 1 if case == "p1_2d" else 2
    dofs, n_dofs, _ = (cells, nodes.shape[0], None) if R == 1 else orc.enumerate_dofs(R, nodes.shape[0], cells, bnd)
    mesh = fdb.Triangulation(nodes, cells, bnd)
    M = cells.shape[1] - 1
    sym_op = -fdb.laplacian() + fdb.reaction(0.7)
    gen_op = -fdb.laplacian() + fdb.advection([1.0, -0.5, 0.25][:nodes.shape[1]]) + fdb.reaction(2.0)

    def run():
        out = []
        s = fdb.Space(mesh, R, dofs, n_dofs)
        for op in (sym_op, gen_op):
            A = fdb.Matrix(s)
            first = A.assemble(op).download_csc()
            s.prepare(op.is_symmetric)
            second = A.assemble(op).download_csc()
            out.append((first, second, s.last_path()[0]))
        return out

    rows = run()
    monkeypatch.setenv("FDB_PATTERN_SORT", "1")
    sort = run()
    for (a1, a2, pa), (b1, b2, pb) in zip(rows, sort):
        assert pa == pb
        for x, y in ((a1, b1), (a2, b2)):
            assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]), "pattern differs between the two builds"
            assert x[2].tobytes() == y[2].tobytes(), "values differ between the two builds"
    # and against the oracle's setFromTriplets pattern
    o_ref, i_ref, _ = orc.assemble_operator(R, nodes, cells, dofs, n_dofs, [(orc.LAPLACIAN, -1.0)], True)
    assert np.array_equal(rows[0][0][0], o_ref) and np.array_equal(rows[0][0][1], i_ref)


@pytest.mark.parametrize("case", ["p1_3d", "p1_2d", "p1_2d_mass", "p2_2d_nonsym", "p2_2d_stiff", "p2_2d_mass", "p1_3d_generic",
                                  "p2_3d", "p2_3d_nonsym", "p2_3d_stiff", "p2_3d_mass", "p1_3d_sphere", "p1_3d_sphere_adr"])
def test_fused_and_two_kernel_paths_are_bit_identical(fdb, golden_meshes, case):
    # the fused path (local matrices in shared memory) and the contribution-list path sum every entry in the same order
    if case in ("p1_3d", "p1_3d_generic"):
        nodes, cells, bnd = fdb.meshes.unit_cube(14)
        nodes = fdb.meshes.jitter(nodes, bnd, 1.0 / 14)
        R, dofs, n_dofs = 1, cells, nodes.shape[0]
        expr = -fdb.laplacian() if case == "p1_3d" else -fdb.diffusion(np.diag([1.0, 2.0, 3.0])) + fdb.reaction(0.5)
    elif case in ("p1_3d_sphere", "p1_3d_sphere_adr"):   # the reference's unstructured ball: ragged row blocks and node lists
        nodes, cells, bnd = golden_meshes("unit_sphere")
        R, dofs, n_dofs = 1, cells, nodes.shape[0]
        expr = (-fdb.laplacian() if case == "p1_3d_sphere"
                else -fdb.laplacian() + fdb.advection([1.0, -0.5, 0.25]) + fdb.reaction(2.0))
    elif case in ("p1_2d", "p1_2d_mass"):
        nodes, cells, bnd = golden_meshes("unit_square")
        R, dofs, n_dofs, expr = 1, cells, nodes.shape[0], (-fdb.laplacian() if case == "p1_2d" else fdb.reaction(1.0))
    elif case.startswith("p2_3d"):   # extension A10: reference-tensor kernels, fused and contribution-list
        nodes, cells, bnd = fdb.meshes.unit_cube(6)
        nodes = fdb.meshes.jitter(nodes, bnd, 1.0 / 6)
        dofs, n_dofs, _ = orc.enumerate_dofs(2, nodes.shape[0], cells, bnd)
        R = 2
        expr = {"p2_3d": -fdb.laplacian() + fdb.reaction(1.5),                     # every tensor (generic rows)
                "p2_3d_stiff": -fdb.laplacian(),                                   # symmetrised Laplacian rows
                "p2_3d_mass": fdb.reaction(1.0),                                   # reaction rows
                "p2_3d_nonsym": -fdb.diffusion([[2.0, 0.3, 0.0], [0.3, 1.0, 0.1], [0.0, 0.1, 1.5]])
                                + fdb.advection([1.0, 0.0, -1.0])}[case]
    else:
        nodes, cells, bnd = golden_meshes("unit_square")
        dofs, n_dofs, _ = orc.enumerate_dofs(2, nodes.shape[0], cells, bnd)
        R = 2
        expr = {"p2_2d_nonsym": -fdb.laplacian() + fdb.advection([1.0, -0.5]) + fdb.reaction(2.0),
                "p2_2d_stiff": -fdb.laplacian(), "p2_2d_mass": fdb.reaction(1.0)}[case]
    mesh = fdb.Triangulation(nodes, cells, bnd)
    s = fdb.Space(mesh, R, dofs, n_dofs)
    A = fdb.Matrix(s)
    two_first = A.assemble(expr).download_csc()   # a first assembly runs the two-kernel path (the plan is lazy)
    s.prepare(expr.is_symmetric)                  # builds the fused plan
    fused = A.assemble(expr).download_csc()
    if case != "p2_3d_nonsym":   # 100 emission slots per cell: contribution-list path only
        assert s.last_path()[0] == 1, "the fused kernel did not run"
    if case in ("p1_3d", "p1_3d_sphere", "p1_3d_sphere_adr", "p1_2d", "p1_2d_mass", "p2_2d_nonsym", "p2_2d_stiff", "p2_2d_mass"):
        # P1 elements (stiffness / mass / general rows) and P2 triangles take the persistent bulk-copy pipeline by default
        assert s.last_kernel() == 2, "the persistent fused kernel did not run"
    assert fused[2].tobytes() == two_first[2].tobytes()
    s.set_fused(False)
    two = A.assemble(expr).download_csc()
    assert fused[2].tobytes() == two[2].tobytes()
    assert np.array_equal(fused[0], two[0]) and np.array_equal(fused[1], two[1])
    assert s.last_path()[0] == 0
    if case.startswith("p2_") or case.endswith("_mass"):   # the specialised reference-tensor rows against the oracle
        M = nodes.shape[1]
        o, i, v = orc.assemble_operator(R, nodes, cells, dofs, n_dofs, orc_terms(expr, M), expr.is_symmetric)
        assert np.array_equal(fused[0], o) and np.array_equal(fused[1], i)
        assert not (np.abs(fused[2] - v) > entry_tolerance(o, i, v, ENTRY_RTOL)).any()


def test_pass_cells_separately(fdb, golden_meshes):
    pts, els, bnd = golden_meshes("c_shaped")
    mesh = fdb.Triangulation(pts, els, bnd)
    s = fdb.Space(mesh, 1, els, pts.shape[0], pass_cells=True)
    v1 = fdb.Matrix(s).assemble(-fdb.laplacian()).download_csc()[2]
    v2 = fdb.Assembler(mesh, 1, pts.shape[0], els).discretize_operator(-fdb.laplacian())[2]
    assert v1.tobytes() == v2.tobytes()


@pytest.mark.parametrize("mesh,R", [("unit_square", 2), ("unit_sphere", 2), ("c_shaped", 1)])
def test_renumbered_dof_table_with_explicit_cells(fdb, golden_meshes, mesh, R):
    """The local problems of the multi-GPU P2 path renumber the dofs (owned first, halo after) and the mesh nodes
    independently, so vertex dofs no longer equal node ids: geometry must come from the cells, indices from the dof
    table.  Checked against the oracle on the same permuted tables, for the matrix, dof coordinates and load vector."""
    pts, els, bnd = golden_meshes(mesh)
    dofs, n_dofs, _ = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
    rng = np.random.default_rng(11)
    perm = rng.permutation(n_dofs).astype(np.int32)
    pdofs = np.asfortranarray(perm[dofs])
    m = fdb.Triangulation(pts, els, bnd)
    s = fdb.Space(m, R, pdofs, n_dofs, pass_cells=True)
    expr = -fdb.laplacian() + fdb.reaction(2.0)
    outer, inner, val = fdb.Matrix(s).assemble(expr).download_csc()
    o, i, v = orc.assemble_operator(R, pts, els, pdofs, n_dofs, orc_terms(expr, pts.shape[1]), True)
    assert np.array_equal(outer, o) and np.array_equal(inner, i)
    assert not (np.abs(val - v) > entry_tolerance(o, i, v, ENTRY_RTOL)).any()
    xo = orc.dofs_coords(R, pts, els, dofs, n_dofs)          # reference numbering
    xp = np.empty_like(xo)
    xp[perm] = xo
    assert np.allclose(s.dofs_coords(), xp, rtol=0, atol=1e-15)   # dof perm[d] sits where dof d sat
    # world = 1 partition of the same space is the identity
    loc = fdb.partition.partition_dofs(pts, els, dofs, n_dofs, np.zeros(n_dofs, np.uint8), 0, 1)
    assert loc.n_owned == n_dofs and np.array_equal(loc.dofs, dofs) and len(loc.neighbors) == 0


# ---- A9: load vector, quadrature nodes, dof coordinates ---------------------------------------------------------

@pytest.mark.parametrize("mesh,R", [("unit_square", 1), ("unit_square", 2), ("unit_sphere", 1), ("unit_sphere", 2),
                                    ("surface", 1), ("surface", 2)])
def test_forcing_quadrature_nodes_and_dof_coords(fdb, golden_meshes, mesh, R):
    pts, els, bnd = golden_meshes(mesh)
    dofs, n_dofs, _ = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
    m = fdb.Triangulation(pts, els, bnd)
    asm = fdb.Assembler(m, R, n_dofs, dofs)
    q = asm.space.quadrature_nodes()
    q_ref = orc.quadrature_nodes(R, pts, els)
    assert np.max(np.abs(q - q_ref)) < 1e-15
    f = np.sin(3 * q_ref[:, 0]) + q_ref[:, -1] ** 2
    b = asm.discretize_forcing(f)
    b_ref = orc.assemble_forcing(R, pts, els, dofs, n_dofs, f)
    assert np.max(np.abs(b - b_ref)) <= 1e-12 * np.abs(b_ref).max()
    xy = asm.space.dofs_coords()
    assert np.max(np.abs(xy - orc.dofs_coords(R, pts, els, dofs, n_dofs))) < 1e-15


# ---- A2/A3: dof enumeration on the device -------------------------------------------------------------------------

@pytest.mark.parametrize("mesh", ["c_shaped", "unit_square", "unit_sphere"])
def test_enumerate_dofs_matches_reference_numbering(fdb, golden_meshes, mesh):
    pts, els, bnd = golden_meshes(mesh)
    m = fdb.Triangulation(pts, els, bnd)
    for R in (1, 2):
        basis = fdb.LagrangianBasis(m, R)
        dofs, n_dofs, bd = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
        assert basis.size() == n_dofs
        assert np.array_equal(basis.dofs(), dofs)
        assert np.array_equal(basis.boundary_dofs(), bd)
    if mesh == "unit_square":
        assert fdb.LagrangianBasis(m, 2).size() == 14161  # 3600 + 10561 (mesh_loader.h:35)


def test_enumerate_dofs_kuhn_cube(fdb):
    nodes, cells, bnd = fdb.meshes.unit_cube(6)
    basis = fdb.LagrangianBasis(fdb.Triangulation(nodes, cells, bnd), 2)
    dofs, n_dofs, bd = orc.enumerate_dofs(2, nodes.shape[0], cells, bnd)
    assert basis.size() == n_dofs and np.array_equal(basis.dofs(), dofs) and np.array_equal(basis.boundary_dofs(), bd)


# ---- A9c/A9d: Dirichlet rows + solve -------------------------------------------------------------------------------

def lu_reference(R, pts, els, bnd, expr, f_fn, g_fn):
    dofs, n_dofs, bd = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
    o, i, v = orc.assemble_operator(R, pts, els, dofs, n_dofs, orc_terms(expr, pts.shape[1]), expr.is_symmetric)
    q = orc.quadrature_nodes(R, pts, els)
    b = orc.assemble_forcing(R, pts, els, dofs, n_dofs, f_fn(q))
    xy = orc.dofs_coords(R, pts, els, dofs, n_dofs)
    g = g_fn(xy)
    orc.set_dirichlet(o, i, v, bd, g, b)
    u = spla.splu(sp.csc_matrix((v, i, o), shape=(n_dofs, n_dofs)), permc_spec="COLAMD").solve(b)
    return u, (o, i, v), b, xy


def test_dirichlet_rows_match_reference_semantics(fdb, golden_meshes):
    pts, els, bnd = golden_meshes("unit_square")
    expr = -fdb.laplacian()
    u, (o, i, v), b_ref, xy = lu_reference(1, pts, els, bnd, expr, lambda q: np.ones(q.shape[0]),
                                           lambda x: x[:, 0] + x[:, 1])
    pde = fdb.PDE(fdb.Triangulation(pts, els, bnd), expr, 1, forcing=lambda q: np.ones(q.shape[0]),
                  solver=fdb.SolverOptions("cg", rtol=1e-12))
    pde.set_dirichlet_bc(xy[:, 0] + xy[:, 1])
    pde.init()
    pde.solve()
    outer, inner, val = pde.stiff()
    assert np.array_equal(outer, o) and np.array_equal(inner, i)  # zeros of replaced rows stay in the pattern
    assert np.all(np.abs(val - v) <= entry_tolerance(o, i, v, ENTRY_RTOL))
    assert np.max(np.abs(pde.force() - b_ref)) <= 1e-12 * np.abs(b_ref).max()
    assert pde.success
    assert np.linalg.norm(pde.solution() - u) / np.linalg.norm(u) < SOLUTION_RTOL


def test_dirichlet_dof0_quirk_with_interior_dof0(fdb, golden_meshes):
    """fem_solver_base.h:86: boundary_dofs_begin() returns index 0 without testing the flag, so dof 0 is ALWAYS treated
    as a Dirichlet dof.  Every stock mesh has node 0 on the boundary; here the nodes are renumbered so that dof 0 is an
    interior node, which makes the quirk visible: row 0 must become a unit row with b(0) = g(0)."""
    pts, els, bnd = golden_meshes("unit_square")
    bnd = np.asarray(bnd).ravel()
    k = int(np.nonzero(bnd == 0)[0][len(bnd) // 3])          # some interior node
    perm = np.arange(pts.shape[0])
    perm[0], perm[k] = k, 0                                   # new id -> old id (swap 0 and k)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    pts2, bnd2, els2 = pts[perm], bnd[perm], inv[els].astype(np.int32)
    assert bnd2[0] == 0
    expr = -fdb.laplacian()
    g = lambda x: x[:, 0] + 2 * x[:, 1]
    u, (o, i, v), b_ref, xy = lu_reference(1, pts2, els2, bnd2, expr, lambda q: np.ones(q.shape[0]), g)
    # the oracle mirrors the quirk: row 0 is a unit row although dof 0 is not flagged
    csr = sp.csc_matrix((v, i, o), shape=(pts.shape[0],) * 2).tocsr()
    assert csr[0].nnz > 1 and np.count_nonzero(csr[0].data) == 1 and csr[0, 0] == 1.0 and b_ref[0] == g(xy)[0]
    pde = fdb.PDE(fdb.Triangulation(pts2, els2, bnd2), expr, 1, forcing=lambda q: np.ones(q.shape[0]),
                  solver=fdb.SolverOptions("cg", rtol=1e-12))
    pde.set_dirichlet_bc(g(xy))
    pde.init()
    pde.solve()
    outer, inner, val = pde.stiff()
    assert np.array_equal(outer, o) and np.array_equal(inner, i)
    assert np.all(np.abs(val - v) <= entry_tolerance(o, i, v, ENTRY_RTOL))
    assert pde.force()[0] == b_ref[0] and pde.solution()[0] == g(xy)[0]
    assert np.linalg.norm(pde.solution() - u) / np.linalg.norm(u) < SOLUTION_RTOL
    # and with the rule switched off (a rank that does not own global dof 0) row 0 stays an ordinary row
    s = fdb.Space(fdb.Triangulation(pts2, els2, bnd2), 1, els2, pts.shape[0], bnd2)
    s.set_dof0_rule(False)
    A = fdb.Matrix(s).assemble(expr)
    bb, xx = fdb.Vector(pts.shape[0]).fill(0.0), fdb.Vector(pts.shape[0]).fill(0.0)
    A.set_dirichlet(fdb.Vector(pts.shape[0], g(xy)), bb, xx)
    o2, i2, v2 = A.download_csc()
    row0 = sp.csc_matrix((v2, i2, o2), shape=(pts.shape[0],) * 2).tocsr()[0]
    assert np.count_nonzero(row0.data) > 1


@pytest.mark.parametrize("R", [1, 2])
def test_fem_pde_laplace_cases(fdb, golden_meshes, R):
    # fem_pde_test.cpp:43-75 (P1, u = x + y, f = 0) and :78-107 (P2, u = 1 - x^2 - y^2, f = 4), threshold 1e-7
    pts, els, bnd = golden_meshes("unit_square")
    ex = (lambda x: x[:, 0] + x[:, 1]) if R == 1 else (lambda x: 1.0 - x[:, 0] ** 2 - x[:, 1] ** 2)
    fv = 0.0 if R == 1 else 4.0
    pde = fdb.PDE(fdb.Triangulation(pts, els, bnd), -fdb.laplacian(), R, forcing=lambda q: np.full(q.shape[0], fv),
                  solver=fdb.SolverOptions("cg", rtol=1e-12))
    xy = pde.dof_coords()
    pde.set_dirichlet_bc(ex(xy))
    pde.init()
    pde.solve()
    assert pde.success
    mo, mi, mv = pde.mass()
    err = ex(xy) - pde.solution()
    assert float((sp.csc_matrix((mv, mi, mo)) @ (err * err)).sum()) < 1e-7
    u, *_ = lu_reference(R, pts, els, bnd, -fdb.laplacian(), lambda q: np.full(q.shape[0], fv), ex)
    assert np.linalg.norm(pde.solution() - u) / np.linalg.norm(u) < SOLUTION_RTOL


@pytest.mark.parametrize("R,tol,jacobi", [(1, 1e-5, False), (2, 1e-7, False), (2, 1e-7, True)])
def test_fem_pde_advection_diffusion_bicgstab(fdb, golden_meshes, R, tol, jacobi):
    # fem_pde_test.cpp:113-166 and :172-212
    from test_oracle_golden import _advdiff_exact
    pts, els, bnd = golden_meshes("unit_square")
    ex, f = _advdiff_exact()
    expr = -fdb.laplacian() + fdb.advection([-1.0, 0.0])
    pde = fdb.PDE(fdb.Triangulation(pts, els, bnd), expr, R, forcing=f,
                  solver=fdb.SolverOptions("bicgstab", rtol=1e-12, jacobi=jacobi))
    pde.set_dirichlet_bc(np.zeros(pde.n_dofs()))
    pde.init()
    pde.solve()
    assert pde.success
    xy = pde.dof_coords()
    mo, mi, mv = pde.mass()
    err = ex(xy) - pde.solution()
    assert float((sp.csc_matrix((mv, mi, mo)) @ (err * err)).sum()) < tol
    u, *_ = lu_reference(R, pts, els, bnd, expr, f, lambda x: np.zeros(x.shape[0]))
    assert np.linalg.norm(pde.solution() - u) / np.linalg.norm(u) < SOLUTION_RTOL


@pytest.mark.parametrize("n,jacobi", [(8, False), (16, False), (16, True), (32, False)])
def test_cg_3d_ladder_vs_lu(fdb, n, jacobi):
    nodes, cells, bnd = fdb.meshes.unit_cube(n)
    f = lambda q: 3 * np.pi ** 2 * np.prod(np.sin(np.pi * q), axis=1)
    pde = fdb.PDE(fdb.Triangulation(nodes, cells, bnd), -fdb.laplacian(), 1, forcing=f,
                  solver=fdb.SolverOptions("cg", rtol=1e-10, jacobi=jacobi))
    pde.set_dirichlet_bc(np.zeros(nodes.shape[0]))
    pde.init()
    pde.solve()
    assert pde.success and pde.stats["rel_resid"] <= 1e-10
    u, *_ = lu_reference(1, nodes, cells, bnd, -fdb.laplacian(), f, lambda x: np.zeros(x.shape[0]))
    assert np.linalg.norm(pde.solution() - u) / np.linalg.norm(u) < SOLUTION_RTOL


def test_spmv_and_true_residual(fdb):
    nodes, cells, bnd = fdb.meshes.unit_cube(10)
    n = nodes.shape[0]
    s = fdb.Space(fdb.Triangulation(nodes, cells, bnd), 1, cells, n, bnd)
    A = fdb.Matrix(s).assemble(-fdb.laplacian() + fdb.reaction(1.0))
    o, i, v = A.download_csc()
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n)
    y = fdb.Vector(n)
    A.spmv(fdb.Vector(n, x), y)
    y_ref = sp.csc_matrix((v, i, o), shape=(n, n)) @ x
    assert np.max(np.abs(y.download() - y_ref)) <= 1e-13 * np.abs(y_ref).max()
    xs, st = A.solve_host(y_ref, np.zeros(n), fdb.SolverOptions("cg", rtol=1e-12))
    assert st["converged"] and np.linalg.norm(xs - x) / np.linalg.norm(x) < 1e-9


def test_solver_repeatable_and_iteration_count_matches_cpu_cg(fdb):
    nodes, cells, bnd = fdb.meshes.unit_cube(12)
    n = nodes.shape[0]
    f = 3 * np.pi ** 2 * np.prod(np.sin(np.pi * orc.quadrature_nodes(1, nodes, cells)), axis=1)
    o, i, v = orc.assemble_operator(1, nodes, cells, cells, n, [(orc.LAPLACIAN, -1.0)], True)
    b = orc.assemble_forcing(1, nodes, cells, cells, n, f)
    orc.set_dirichlet(o, i, v, bnd, np.zeros(n), b)
    Ar = sp.csc_matrix((v, i, o), shape=(n, n)).tocsr()
    Ar.sort_indices()
    u_cpu, it_cpu, _ = orc.cg(Ar.indptr, Ar.indices, Ar.data, b, np.zeros(n), rtol=1e-8)
    runs = []
    for _ in range(2):
        pde = fdb.PDE(fdb.Triangulation(nodes, cells, bnd), -fdb.laplacian(), 1, forcing=lambda q: f,
                      solver=fdb.SolverOptions("cg", rtol=1e-8, check_every=7))
        pde.set_dirichlet_bc(np.zeros(n))
        pde.init()
        pde.solve()
        runs.append((pde.solution().tobytes(), pde.stats["iters"]))
    assert runs[0] == runs[1]
    assert abs(runs[0][1] - it_cpu) <= 2
    assert np.linalg.norm(np.frombuffer(runs[0][0]) - u_cpu) / np.linalg.norm(u_cpu) < 1e-7


# ---- error behaviour (utils/assert.h:23-27, fem_linear_elliptic_solver.h:36) --------------------------------------

def test_error_behaviour(fdb):
    nodes, cells, bnd = fdb.meshes.unit_square(4)
    s = fdb.Space(fdb.Triangulation(nodes, cells, bnd), 1, cells, nodes.shape[0])
    A = fdb.Matrix(s)
    with pytest.raises(fdb.FdbError) as ei:
        A.solve(fdb.Vector(25), fdb.Vector(25), fdb.SolverOptions())
    assert ei.value.code == 3 and "initialized" in str(ei.value)
    A.assemble(-fdb.laplacian())
    with pytest.raises(fdb.FdbError) as ei:  # boundary markers never uploaded
        A.set_dirichlet(fdb.Vector(25), fdb.Vector(25))
    assert ei.value.code == 3
    # iteration budget exhausted: `success = false` (fem_linear_elliptic_solver.h:42-45), no throw across the ABI
    A.assemble(-fdb.laplacian() + fdb.reaction(1.0))
    rhs = np.random.default_rng(0).standard_normal(25)
    st = A.solve(fdb.Vector(25, rhs), fdb.Vector(25).fill(0.0), fdb.SolverOptions("cg", maxit=2, rtol=1e-14),
                 raise_on_fail=False)
    assert not st["converged"] and st["iters"] == 2
    with pytest.raises(fdb.FdbError) as ei:
        A.solve(fdb.Vector(25, rhs), fdb.Vector(25).fill(0.0), fdb.SolverOptions("cg", maxit=2, rtol=1e-14))
    assert ei.value.code == 4
    # zero right-hand side -> zero solution
    x = fdb.Vector(25, np.ones(25))
    st = A.solve(fdb.Vector(25).fill(0.0), x, fdb.SolverOptions("cg"))
    assert st["converged"] and np.all(x.download() == 0.0)


def test_persistent_cg_kernel_matches_multi_kernel_loop(fdb):
    # the cooperative single-kernel CG runs the same recurrences as the multi-kernel loop
    nodes, cells, bnd = fdb.meshes.unit_cube(14)
    n = nodes.shape[0]
    f = lambda q: 3 * np.pi ** 2 * np.prod(np.sin(np.pi * q), axis=1)
    sols = []
    try:
        for mode, jac in ((1, False), (2, False), (1, True), (2, True)):
            assert fdb.lib().fdb_set_persistent_cg(mode) == 0
            pde = fdb.PDE(fdb.Triangulation(nodes, cells, bnd), -fdb.laplacian(), 1, forcing=f,
                          solver=fdb.SolverOptions("cg", rtol=1e-11, jacobi=jac))
            pde.set_dirichlet_bc(np.zeros(n))
            pde.init()
            pde.solve()
            assert pde.success
            sols.append((pde.solution(), pde.stats["iters"]))
    finally:
        fdb.lib().fdb_set_persistent_cg(1)
    for a, b in ((0, 1), (2, 3)):
        assert abs(sols[a][1] - sols[b][1]) <= 1
        assert np.linalg.norm(sols[a][0] - sols[b][0]) / np.linalg.norm(sols[a][0]) < 1e-9


# ---- N2: parabolic driver (fem_linear_parabolic_solver.h:37-72) ------------------------------------------------------

from parabolic_ref import parabolic_reference as _parabolic_reference  # noqa: E402


def test_parabolic_isotropic_order1_convergence(fdb):
    # fem_pde_test.cpp:295-368 through fdb_solve_parabolic: P1, 31 time steps on unit_square_{16,32,64,128}; the L2 error at
    # the final time falls with order 2 (floor(log2(e_k / e_{k+1})) == 2), and every mesh agrees with the oracle time loop
    pi = np.pi
    times = np.linspace(0.0, 1.0, 31)
    u_fn = lambda x, t: np.sin(2 * pi * x[:, 0]) * np.sin(2 * pi * x[:, 1]) * np.exp(-t)
    f_fn = lambda x, t: (8 * pi * pi - 1.0) * np.sin(2 * pi * x[:, 0]) * np.sin(2 * pi * x[:, 1]) * np.exp(-t)
    errs = []
    for N in (16, 32, 64, 128):
        pts, els, bnd = fdb.meshes.unit_square(N)
        n = pts.shape[0]
        s = fdb.Space(fdb.Triangulation(pts, els, bnd), 1, els, n, bnd)
        stiff = fdb.Matrix(s).assemble(fdb.dt() - fdb.laplacian())
        mass = fdb.Matrix(s).assemble(fdb.reaction(1.0))
        xy, q = s.dofs_coords(), s.quadrature_nodes()
        f = np.stack([f_fn(q, t) for t in times], axis=1)
        g = np.stack([u_fn(xy, t) for t in times], axis=1)
        sol, st = fdb.solve_parabolic(stiff, mass, times[1] - times[0], f, g, u_fn(xy, times[0]),
                                      fdb.SolverOptions("cg", rtol=1e-12))
        assert st["converged"]
        mo, mi, mv = mass.download_csc()
        Mass = sp.csc_matrix((mv, mi, mo), shape=(n, n))
        e = g[:, -1] - sol[:, -1]
        errs.append(np.sqrt(float((Mass @ (e * e)).sum())))
        if N <= 32:   # the oracle's SuperLU time loop on the same mesh
            ref, _, _, _ = _parabolic_reference(1, pts, els, bnd, times, u_fn, f_fn)
            assert np.linalg.norm(sol[:, -1] - ref[:, -1]) / np.linalg.norm(ref[:, -1]) < SOLUTION_RTOL
    orders = [np.log2(errs[k] / errs[k + 1]) for k in range(3)]
    assert all(np.floor(o) == 2 for o in orders), (errs, orders)


def test_parabolic_isotropic_order2(fdb, golden_meshes):
    # fem_pde_test.cpp:222-285: dt(u) - lap u = f on unit_square, P2, 101 time steps, error (mass*err^2).sum() < 1e-7
    pts, els, bnd = golden_meshes("unit_square")
    pi = np.pi
    times = np.linspace(0.0, 1.0, 101)
    u_fn = lambda x, t: np.sin(2 * pi * x[:, 0]) * np.sin(2 * pi * x[:, 1]) * np.exp(-t)
    f_fn = lambda x, t: (8 * pi * pi - 1.0) * np.sin(2 * pi * x[:, 0]) * np.sin(2 * pi * x[:, 1]) * np.exp(-t)
    mesh = fdb.Triangulation(pts, els, bnd)
    basis = fdb.LagrangianBasis(mesh, 2)
    n = basis.size()
    s = fdb.Space(mesh, 2, basis.dofs(), n, basis.boundary_dofs())
    L = fdb.dt() - fdb.laplacian()
    stiff = fdb.Matrix(s).assemble(L)
    mass = fdb.Matrix(s).assemble(fdb.reaction(1.0))
    xy, q = s.dofs_coords(), s.quadrature_nodes()
    f = np.stack([f_fn(q, t) for t in times], axis=1)
    g = np.stack([u_fn(xy, t) for t in times], axis=1)
    sol, st = fdb.solve_parabolic(stiff, mass, times[1] - times[0], f, g, u_fn(xy, times[0]),
                                  fdb.SolverOptions("cg", rtol=1e-12))
    assert st["converged"]
    mo, mi, mv = mass.download_csc()
    Mass = sp.csc_matrix((mv, mi, mo), shape=(n, n))
    errs = [float((Mass @ ((g[:, j] - sol[:, j]) ** 2)).sum()) for j in range(times.size)]
    assert max(errs) < 1e-7                                            # the reference's own acceptance threshold
    ref, xy_ref, _, _ = _parabolic_reference(2, pts, els, bnd, times, u_fn, f_fn)
    assert np.max(np.abs(xy - xy_ref)) < 1e-15
    for j in (1, 10, 50, 100):
        assert np.linalg.norm(sol[:, j] - ref[:, j]) / np.linalg.norm(ref[:, j]) < SOLUTION_RTOL
