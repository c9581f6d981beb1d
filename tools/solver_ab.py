"""Development helper: the Krylov solver variants side by side on one GPU (C4: CG, C3: BiCGSTAB).
   gpurun -- python tools/solver_ab.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
fdb = g.load_package()
L = fdb.lib()
which = os.environ.get("AB_SOLVERS", "c4,c3")

def run(A, b, n, kind, label, maxit=0):
    out = []
    x = fdb.Vector(n)
    for rep in range(3):
        x.fill(0.0)
        A.set_dirichlet(fdb.Vector(n).fill(0.0), b, x)
        st = A.solve(b, x, fdb.SolverOptions(kind, rtol=1e-8, maxit=maxit, check_every=50), raise_on_fail=False)
        out.append(st)
    st = out[-1]
    print(f"{label:28s} iters {st['iters']:5d} conv {st['converged']} resid {st['rel_resid']:.3e} "
          f"{st['seconds']*1e3:9.3f} ms  {st['seconds']/max(st['iters'],1)*1e6:7.2f} us/iter", flush=True)
    return x.download()

if "c4" in which:
    nodes, cells, bnd = fdb.meshes.unit_cube(int(os.environ.get("AB_N", "119")))
    n = nodes.shape[0]
    s = fdb.Space(fdb.Triangulation(nodes, cells, bnd), 1, cells, n, bnd)
    A = fdb.Matrix(s).assemble(-fdb.laplacian())
    q = s.quadrature_nodes()
    f = 3 * np.pi ** 2 * np.prod(np.sin(np.pi * q), axis=1)
    b, fq = fdb.Vector(n), fdb.Vector(f.size, f)
    assert L.fdb_assemble_forcing(s.h, fq.h, b.h) == 0
    L.fdb_set_persistent_sell(0); L.fdb_set_persistent_cg(0)
    u0 = run(A, b, n, "cg", "C4 CG multi-kernel graph")
    L.fdb_set_persistent_cg(2)
    u1 = run(A, b, n, "cg", "C4 CG persistent (CSR)")
    L.fdb_set_persistent_cg(0); L.fdb_set_persistent_sell(2)
    u2 = run(A, b, n, "cg", "C4 CG1 persistent SELL")
    print("   rel diff persistent-SELL vs multi-kernel:", np.linalg.norm(u2 - u0) / np.linalg.norm(u0),
          " old persistent:", np.linalg.norm(u1 - u0) / np.linalg.norm(u0))
    u3 = run(A, b, n, "bicgstab", "C4 BiCGSTAB persistent SELL")
    L.fdb_set_persistent_sell(0)
    u4 = run(A, b, n, "bicgstab", "C4 BiCGSTAB multi-kernel")
    print("   rel diff BiCGSTAB persistent vs multi-kernel:", np.linalg.norm(u3 - u4) / np.linalg.norm(u4),
          " vs CG:", np.linalg.norm(u3 - u0) / np.linalg.norm(u0))
    del A, s
if "c3" in which:
    nodes, cells, bnd = fdb.meshes.unit_square(1000)
    mesh = fdb.Triangulation(nodes, cells, bnd)
    basis = fdb.LagrangianBasis(mesh, 2)
    n = basis.size()
    s = fdb.Space(mesh, 2, basis.dofs(), n, basis.boundary_dofs())
    A = fdb.Matrix(s).assemble(-fdb.laplacian() + fdb.advection([-1.0, 0.0]) + fdb.reaction(1.0))
    q = s.quadrature_nodes()
    pi = np.pi
    f = (2 * pi ** 2 + 1) * np.sin(pi * q[:, 0]) * np.sin(pi * q[:, 1]) - pi * np.cos(pi * q[:, 0]) * np.sin(pi * q[:, 1])
    b, fq = fdb.Vector(n), fdb.Vector(f.size, f)
    assert L.fdb_assemble_forcing(s.h, fq.h, b.h) == 0
    L.fdb_set_persistent_sell(0)
    u0 = run(A, b, n, "bicgstab", "C3 BiCGSTAB multi-kernel", 30000)
    L.fdb_set_persistent_sell(2)
    u1 = run(A, b, n, "bicgstab", "C3 BiCGSTAB persistent SELL", 30000)
    print("   rel diff:", np.linalg.norm(u1 - u0) / np.linalg.norm(u0))
