"""BASELINE.json configurations at FULL size, checked through size-independent properties (the CPU oracle would need
minutes per case): conservation identities of the assembled operators, bit-reproducibility, symmetry, the true residual
of the Krylov solution and the discretisation error of a manufactured solution."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _csr(outer, inner, val, n):
    return sp.csc_matrix((val, inner, outer), shape=(n, n)).tocsr()


def test_c4_3d_p1_laplacian_10m_tets(fdb):
    # configs[3]: 3D Laplacian P1, unit-cube Kuhn mesh n=119 (10,110,954 tets), assembly + CG to 1e-8
    n_cube = 119
    nodes, cells, bnd = fdb.meshes.unit_cube(n_cube)
    n = nodes.shape[0]
    assert cells.shape[0] == 10110954 and n == 1728000
    mesh = fdb.Triangulation(nodes, cells, bnd)
    s = fdb.Space(mesh, 1, cells, n, bnd)
    A = fdb.Matrix(s).assemble(-fdb.laplacian())
    o, i, v = A.download_csc()
    assert i.size == 25575838
    K = _csr(o, i, v, n)
    assert abs(K - K.T).max() == 0.0                                   # mirrored bit-exactly
    scale = np.abs(v).max()
    assert np.abs(K @ np.ones(n)).max() < 1e-12 * scale * 30           # constants in the kernel
    lin = nodes @ np.array([1.0, -2.0, 0.5])
    interior = bnd == 0
    assert np.abs((K @ lin)[interior]).max() < 1e-11 * scale * 30      # linear functions are discrete-harmonic
    s.prepare(True)                                                    # fused path, same bits
    assert A.assemble(-fdb.laplacian()).download_csc()[2].tobytes() == v.tobytes()
    M = fdb.Matrix(s).assemble(fdb.reaction(1.0))
    mv = M.download_csc()[2]
    assert abs(mv.sum() - 1.0) < 1e-12                                 # sum of the mass matrix = volume
    # manufactured solution u = prod sin(pi x): -lap u = 3 pi^2 u, zero Dirichlet data
    q = s.quadrature_nodes()
    f = 3 * np.pi ** 2 * np.prod(np.sin(np.pi * q), axis=1)
    del q
    b = fdb.Vector(n)
    fq = fdb.Vector(f.size, f)
    assert fdb.lib().fdb_assemble_forcing(s.h, fq.h, b.h) == 0
    x = fdb.Vector(n).fill(0.0)
    A.set_dirichlet(fdb.Vector(n).fill(0.0), b, x)
    st = A.solve(b, x, fdb.SolverOptions("cg", rtol=1e-8))
    assert st["converged"] and st["rel_resid"] <= 1e-8
    u = x.download()
    o2, i2, v2 = A.download_csc()                                       # with Dirichlet rows
    bh = b.download()
    true_res = np.linalg.norm(bh - _csr(o2, i2, v2, n) @ u) / np.linalg.norm(bh)
    assert true_res < 2e-8
    u_ex = np.prod(np.sin(np.pi * nodes), axis=1)
    err = u - u_ex
    l2 = np.sqrt(float(err @ (_csr(*M.download_csc(), n) @ err)))
    assert l2 < 5e-4                                                    # O(h^2), h = 1/119


def test_c2_2d_p1_poisson_4m_triangles(fdb):
    # configs[1]: 2D Poisson P1 on the structured unit square, 4M triangles, stiffness + mass + CG
    N = 1414
    nodes, cells, bnd = fdb.meshes.unit_square(N)
    n = nodes.shape[0]
    assert cells.shape[0] == 3998792 and n == 2002225
    s = fdb.Space(fdb.Triangulation(nodes, cells, bnd), 1, cells, n, bnd)
    A = fdb.Matrix(s).assemble(-fdb.laplacian())
    M = fdb.Matrix(s).assemble(fdb.reaction(1.0))
    o, i, v = A.download_csc()
    assert i.size == 14004257                                          # nodes + 2 edges, explicit zeros kept
    K = _csr(o, i, v, n)
    assert abs(K - K.T).max() == 0.0 and np.abs(K @ np.ones(n)).max() < 1e-11
    mo, mi, mv = M.download_csc()
    assert abs(mv.sum() - 1.0) < 1e-12
    q = s.quadrature_nodes()
    f = 2 * np.pi ** 2 * np.sin(np.pi * q[:, 0]) * np.sin(np.pi * q[:, 1])
    b = fdb.Vector(n)
    fq = fdb.Vector(f.size, f)  # keep the handle alive across the call
    assert fdb.lib().fdb_assemble_forcing(s.h, fq.h, b.h) == 0
    x = fdb.Vector(n).fill(0.0)
    A.set_dirichlet(fdb.Vector(n).fill(0.0), b, x)
    st = A.solve(b, x, fdb.SolverOptions("cg", rtol=1e-8))
    assert st["converged"]
    err = x.download() - np.sin(np.pi * nodes[:, 0]) * np.sin(np.pi * nodes[:, 1])
    assert np.sqrt(float(err @ (_csr(mo, mi, mv, n) @ err))) < 1e-5


def test_c3_2d_p2_advection_diffusion_reaction_2m_triangles(fdb):
    # configs[2]: P2, -laplacian + advection(b = (-1, 0)) + reaction(1), 2M triangles, BiCGSTAB (1 GPU here)
    N = 1000
    nodes, cells, bnd = fdb.meshes.unit_square(N)
    mesh = fdb.Triangulation(nodes, cells, bnd)
    basis = fdb.LagrangianBasis(mesh, 2)
    n = basis.size()
    assert n == 4004001 and cells.shape[0] == 2000000
    s = fdb.Space(mesh, 2, basis.dofs(), n, basis.boundary_dofs())
    L = -fdb.laplacian() + fdb.advection([-1.0, 0.0]) + fdb.reaction(1.0)
    A = fdb.Matrix(s).assemble(L)
    o, i, v = A.download_csc()
    assert i.size == 46016001
    K = _csr(o, i, v, n)
    Mo, Mi, Mv = fdb.Matrix(s).assemble(fdb.reaction(1.0)).download_csc()
    Mass = _csr(Mo, Mi, Mv, n)
    assert abs(Mv.sum() - 1.0) < 1e-11
    # row sums: stiffness and advection annihilate constants, the reaction term leaves the mass row sums
    assert np.abs(K @ np.ones(n) - Mass @ np.ones(n)).max() < 1e-10
    xy = s.dofs_coords()
    q = s.quadrature_nodes()
    pi = np.pi
    # manufactured: u = sin(pi x) sin(pi y); L u = 2 pi^2 u - u_x + u
    f = (2 * pi ** 2 + 1) * np.sin(pi * q[:, 0]) * np.sin(pi * q[:, 1]) - pi * np.cos(pi * q[:, 0]) * np.sin(pi * q[:, 1])
    b = fdb.Vector(n)
    fq = fdb.Vector(f.size, f)  # keep the handle alive across the call
    assert fdb.lib().fdb_assemble_forcing(s.h, fq.h, b.h) == 0
    x = fdb.Vector(n).fill(0.0)
    A.set_dirichlet(fdb.Vector(n).fill(0.0), b, x)
    st = A.solve(b, x, fdb.SolverOptions("bicgstab", rtol=1e-8, maxit=30000))  # 4425 iterations, 1.7 s on B200
    assert st["converged"]
    err = x.download() - np.sin(pi * xy[:, 0]) * np.sin(pi * xy[:, 1])
    assert np.sqrt(float(err @ (Mass @ err))) < 1e-7                   # O(h^3) for P2
