"""BASELINE.json configurations at FULL size, checked through size-independent properties (the CPU oracle would need
minutes per case): conservation identities of the assembled operators, bit-reproducibility, symmetry, the true residual
of the Krylov solution and the discretisation error of a manufactured solution."""
import os

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


def test_c4_full_size_against_the_oracle(fdb):
    """configs[3] at FULL size against the CPU oracle itself: sparsity pattern compared byte for byte, every one of the
    25.6 M entries within 1e-12, and the GPU CG solution against the oracle's CPU CG on the oracle's matrix (same
    stopping rule).  The oracle's all-core assembly is bit-identical to its serial routine
    (tests/test_oracle_golden.py::test_all_core_variant_is_bit_identical); it needs ~6 s here, the CPU CG ~20 s."""
    from conftest import entry_tolerance
    from oracle import oracle as orc
    nodes, cells, bnd = fdb.meshes.unit_cube(119)
    n = nodes.shape[0]
    s = fdb.Space(fdb.Triangulation(nodes, cells, bnd), 1, cells, n, bnd)
    s.prepare(True)
    A = fdb.Matrix(s).assemble(-fdb.laplacian())
    A.assemble(-fdb.laplacian())                                        # the fused kernel (plan prepared above)
    assert s.last_path()[0]
    o, i, v = A.download_csc()
    o_ref, i_ref, v_ref = orc.assemble_operator_mt(1, nodes, cells, cells, n, [(orc.LAPLACIAN, -1.0)], True,
                                                   n_threads=os.cpu_count() or 1)
    assert o.tobytes() == o_ref.tobytes() and i.tobytes() == i_ref.tobytes()      # pattern: memcmp
    tol = entry_tolerance(o_ref, i_ref, v_ref, 1e-12)
    assert not (np.abs(v - v_ref) > tol).any()
    # solve: same right-hand side, same Dirichlet rows, CG to 1e-8 on both sides
    q = orc.quadrature_nodes(1, nodes, cells)
    f = 3 * np.pi ** 2 * np.prod(np.sin(np.pi * q), axis=1)
    del q
    b_ref = orc.assemble_forcing(1, nodes, cells, cells, n, f)
    orc.set_dirichlet(o_ref, i_ref, v_ref, bnd, np.zeros(n), b_ref)
    Ar = _csr(o_ref, i_ref, v_ref, n)
    Ar.sort_indices()
    u_cpu, it_cpu, _ = orc.cg(Ar.indptr, Ar.indices, Ar.data, b_ref, np.zeros(n), rtol=1e-8)
    b, fq = fdb.Vector(n), fdb.Vector(f.size, f)
    assert fdb.lib().fdb_assemble_forcing(s.h, fq.h, b.h) == 0
    x = fdb.Vector(n).fill(0.0)
    A.set_dirichlet(fdb.Vector(n).fill(0.0), b, x)
    assert np.max(np.abs(b.download() - b_ref)) <= 1e-12 * np.abs(b_ref).max()
    st = A.solve(b, x, fdb.SolverOptions("cg", rtol=1e-8))
    assert st["converged"] and abs(st["iters"] - it_cpu) <= 2
    u = x.download()
    assert np.linalg.norm(u - u_cpu) / np.linalg.norm(u_cpu) < 1e-7     # two CG runs stopped at 1e-8 each


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


def test_c5_3d_p2_one_slab_of_eight(fdb):
    """configs[4]: 3D P2 (extension A10), unit-cube Kuhn mesh n=150 (20,250,000 tets, 27.4 M dofs), mass + stiffness
    assembly, 8 GPUs.  Assembly needs no communication, so ONE GPU running rank 3's local problem of the 8-way partition
    measures the per-GPU work and HBM footprint of the 8-GPU job; checked through exact identities."""
    import time
    n_cube, world, rank = 150, 8, 3
    nodes, cells, bnd = fdb.meshes.unit_cube(n_cube)
    assert cells.shape[0] == 20250000
    mesh = fdb.Triangulation(nodes, cells, bnd)
    t0 = time.perf_counter()
    basis = fdb.LagrangianBasis(mesh, 2)                      # global edge numbering on the device
    t_enum = time.perf_counter() - t0
    dofs, nd, bd = basis.dofs(), basis.size(), basis.boundary_dofs()
    t0 = time.perf_counter()
    loc = fdb.partition.partition_dofs(nodes, cells, dofs, nd, bd, rank, world)
    t_part = time.perf_counter() - t0
    del basis, dofs
    lmesh = fdb.Triangulation(loc.nodes, loc.cells, np.zeros(loc.nodes.shape[0], np.uint8))
    s = fdb.Space(lmesh, 2, loc.dofs, loc.n_local_dofs, loc.boundary, pass_cells=True)
    K, Mm = fdb.Matrix(s), fdb.Matrix(s)
    times, split = {}, {}
    s.set_profiling(True)
    s.prepare(True)        # fused plan (row blocks + shared-memory gather lists) for the symmetric pattern
    for name, A, expr in (("stiffness", K, -fdb.laplacian()), ("mass", Mm, fdb.reaction(1.0))):
        A.assemble(expr)
        s.sync()
        t0 = time.perf_counter()
        A.assemble(expr)
        s.sync()
        times[name] = time.perf_counter() - t0
        split[name] = s.last_timings()      # [local kernel, reduction] in ms (CUDA events)
    nl, no = loc.n_local_dofs, loc.n_owned
    ones = fdb.Vector(nl).fill(1.0)
    y = fdb.Vector(nl)
    K.spmv(ones, y)
    ky = y.download()[:no]
    o, i, v = K.download_csc()
    scale = np.abs(v).max()
    assert np.abs(ky).max() < 1e-11 * scale                   # constants are in the kernel of every owned stiffness row
    Mm.spmv(ones, y)
    my = y.download()[:no]
    # (M 1)_i = int psi_i = (cells around i) * |e| * (-1/20 for a vertex dof, 1/5 for an edge dof), |e| = h^3 / 6
    cnt = np.bincount(loc.dofs.ravel(), minlength=nl)[:no]
    is_vertex = loc.local_to_global[:no] < nodes.shape[0]
    vol = (1.0 / n_cube) ** 3 / 6.0
    expect = cnt * vol * np.where(is_vertex, -1.0 / 20.0, 1.0 / 5.0)
    assert np.max(np.abs(my - expect)) < 1e-12 * np.abs(expect).max()
    n_loc_cells = loc.cells.shape[0]
    print(f"[C5 slab {rank}/{world}] local cells {n_loc_cells} ({n_loc_cells / (cells.shape[0] / world):.3f} x share), "
          f"dofs {nl} (owned {no}), nnz {v.size}; enumerate {t_enum:.1f} s, partition {t_part:.1f} s; "
          f"stiffness {times['stiffness'] * 1e3:.2f} ms {split['stiffness']}, mass {times['mass'] * 1e3:.2f} ms {split['mass']} "
          f"-> {cells.shape[0] / world / max(times.values()) / 1e9:.2f} G tets/s per GPU and matrix")
