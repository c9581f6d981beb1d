"""Oracle restatement of the parabolic driver, shared by the CPU pin and the GPU parity test.
FEMSolverBase::init + FEMLinearParabolicSolver::solve (solvers/fem_solver_base.h:106-139,
solvers/fem_linear_parabolic_solver.h:37-72): K = mass/dt + stiff, Dirichlet rows, factor once, then per step
rhs = (mass/dt) u_i + force_{i+1} with the boundary values of time i+1."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import oracle as orc


def parabolic_reference(R, pts, els, bnd, times, u_fn, f_fn):
    """Oracle restatement of FEMSolverBase::init + FEMLinearParabolicSolver::solve with SuperLU (factor once)."""
    dofs, n, bd = orc.enumerate_dofs(R, pts.shape[0], els, bnd)
    xy = orc.dofs_coords(R, pts, els, dofs, n)
    q = orc.quadrature_nodes(R, pts, els)
    so, si, sv = orc.assemble_operator(R, pts, els, dofs, n, [(orc.DT, 1.0), (orc.LAPLACIAN, -1.0)], True)
    mo, mi, mv = orc.assemble_operator(R, pts, els, dofs, n, [(orc.REACTION, 1.0, [1.0])], True)
    dt_ = times[1] - times[0]
    mass = sp.csc_matrix((mv, mi, mo), shape=(n, n))
    ko, ki, kv = so.copy(), si.copy(), mv / dt_ + sv                      # same pattern: K = mass/dt + stiff
    dummy = np.zeros(n)
    orc.set_dirichlet(ko, ki, kv, bd, np.zeros(n), dummy)
    lu = spla.splu(sp.csc_matrix((kv, ki, ko), shape=(n, n)), permc_spec="COLAMD")
    sol = np.zeros((n, times.size))
    sol[:, 0] = u_fn(xy, times[0])
    isb = (bd > 0) | (np.arange(n) == 0)
    for i in range(times.size - 1):
        rhs = (mass @ sol[:, i]) / dt_ + orc.assemble_forcing(R, pts, els, dofs, n, f_fn(q, times[i + 1]))
        rhs[isb] = u_fn(xy, times[i + 1])[isb]
        sol[:, i + 1] = lu.solve(rhs)
    return sol, xy, q, mass
