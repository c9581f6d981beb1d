"""ctypes binding of libfdapde_b200.so + the host-side mirror of the reference's interface for the hot path.

Reference interfaces mirrored (paths relative to the fdaPDE-core tree):
  Triangulation<M,N>              fdaPDE/geometry/triangulation.h:36-125
  LagrangianBasis<Mesh,R>         fdaPDE/finite_elements/basis/lagrangian_basis.h:31,147-184
  Assembler<FEM,D,B,I>            fdaPDE/finite_elements/fem_assembler.h:36-136
  operator expressions            fdaPDE/pde/differential_operators.h:27-52, differential_expressions.h:38-135
  PDE<D,E,F,FEM,fem_order<R>>     fdaPDE/pde/pde.h:40-114 ; FEMSolverBase solvers/fem_solver_base.h:36-155

Nothing here computes: every method marshals numpy arrays into the C ABI (include/fdapde_b200.h).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("FDB_LIB_PATH") or os.path.join(_HERE, "lib", "libfdapde_b200.so")
_lib = None

FDB_OK, FDB_ERR_ARG, FDB_ERR_CUDA, FDB_ERR_STATE, FDB_ERR_NOT_CONVERGED, FDB_ERR_UNSUPPORTED = range(6)
LAPLACIAN, DIFFUSION, ADVECTION, REACTION, DT = range(5)
MAX_TERMS = 8


class FdbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fdapde_b200 error {code}: {msg}")
        self.code = code


class _Term(C.Structure):
    _fields_ = [("kind", C.c_int32), ("space_varying", C.c_int32), ("scale", C.c_double), ("coeff", C.c_void_p)]


class _OpDesc(C.Structure):
    _fields_ = [("n_terms", C.c_int32), ("symmetric", C.c_int32), ("terms", _Term * MAX_TERMS)]


class _SolverOpts(C.Structure):
    _fields_ = [("kind", C.c_int32), ("jacobi", C.c_int32), ("maxit", C.c_int32), ("check_every", C.c_int32),
                ("rtol", C.c_double)]


class _SolveStats(C.Structure):
    _fields_ = [("iters", C.c_int32), ("converged", C.c_int32), ("rel_resid", C.c_double), ("seconds", C.c_double)]


def lib_path():
    return _LIB_PATH


def lib():
    """Loads the CUDA library.  Fails loudly when it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise FdbError(FDB_ERR_CUDA, f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; "
                                         f"g.build()'` (no CPU fallback exists)")
        L = C.CDLL(_LIB_PATH)
        L.fdb_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc):
    if rc != FDB_OK:
        raise FdbError(rc, lib().fdb_last_error().decode())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _PinnedBlock:
    """Owner of one fdb_host_alloc block; numpy arrays built on it keep it alive through their .base chain."""

    def __init__(self, nbytes):
        self.p = C.c_void_p()
        _check(lib().fdb_host_alloc(C.c_size_t(max(int(nbytes), 1)), C.byref(self.p)))
        self.nbytes = int(nbytes)

    def __del__(self):
        if getattr(self, "p", None) and _lib is not None:
            _lib.fdb_host_free(self.p)
            self.p = None


def pinned_empty(shape, dtype=np.float64, order="C"):
    """numpy array in page-locked host memory (fdb_host_alloc): the DMA target / source of the arrays that cross the
    C ABI.  The block returns to the library's cache when the array is garbage collected."""
    dt = np.dtype(dtype)
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(v) for v in shape)
    n = int(np.prod(shape)) if shape else 1
    blk = _PinnedBlock(n * dt.itemsize)
    buf = (C.c_char * max(n * dt.itemsize, 1)).from_address(blk.p.value)
    buf._fdb_block = blk   # ctypes array -> block; numpy keeps the ctypes array as .base
    return np.frombuffer(buf, dtype=dt, count=n).reshape(shape, order=order)


def pinned_copy(a, order="C"):
    a = np.asarray(a)
    out = pinned_empty(a.shape, a.dtype, order=order)
    out[...] = a
    return out


# ---- operator expressions (differential_operators.h / differential_expressions.h) ------------------------------------
class DifferentialExpr:
    """Flattened expression tree: a list of (kind, scale, coeff, space_varying) leaves."""

    def __init__(self, leaves):
        self.leaves = leaves

    @property
    def is_symmetric(self):  # AND over leaves; Advection is not symmetric (advection.h:43)
        return all(k != ADVECTION for k, *_ in self.leaves)

    def __add__(self, o):
        return DifferentialExpr(self.leaves + o.leaves)

    def __sub__(self, o):
        return DifferentialExpr(self.leaves + (-o).leaves)

    def __neg__(self):
        return DifferentialExpr([(k, -s, c, sv) for k, s, c, sv in self.leaves])

    def __rmul__(self, a):
        return DifferentialExpr([(k, float(a) * s, c, sv) for k, s, c, sv in self.leaves])

    def descriptor(self, n_quad_rows=None, symmetric=None):
        # expressions are immutable (every operator returns a new object): the lowered descriptor is built once per
        # (rows, symmetry) and reused by repeated assemblies (time stepping, benchmark loops)
        cache = self.__dict__.setdefault("_desc_cache", {})
        key = (n_quad_rows, symmetric)
        if key in cache:
            return cache[key]
        assert len(self.leaves) <= MAX_TERMS
        d = _OpDesc()
        d.n_terms = len(self.leaves)
        d.symmetric = int(self.is_symmetric if symmetric is None else symmetric)
        keep = []
        for t, (k, s, c, sv) in enumerate(self.leaves):
            d.terms[t].kind = k
            d.terms[t].scale = s
            d.terms[t].space_varying = int(sv)
            if c is not None:
                a = np.ascontiguousarray(c, dtype=np.float64)
                if sv and n_quad_rows is not None:
                    assert a.shape[0] == n_quad_rows, "space-varying coefficient needs n_cells*n_quad rows"
                keep.append(a)
                d.terms[t].coeff = a.ctypes.data
        d._keep = keep
        cache[key] = d
        return d


def laplacian():
    return DifferentialExpr([(LAPLACIAN, 1.0, None, False)])


def dt():
    return DifferentialExpr([(DT, 1.0, None, False)])


def diffusion(K):
    K = np.asarray(K, dtype=np.float64)
    if K.ndim == 2 and K.shape[0] == K.shape[1] and K.shape[0] <= 3:  # constant SMatrix<N>: column-major
        return DifferentialExpr([(DIFFUSION, 1.0, np.asfortranarray(K).ravel(order="F"), False)])
    return DifferentialExpr([(DIFFUSION, 1.0, K, True)])  # rows nq*e+q, each an N*N column-major block


def advection(b):
    b = np.asarray(b, dtype=np.float64)
    return DifferentialExpr([(ADVECTION, 1.0, b, b.ndim == 2)])


def reaction(c):
    c = np.asarray(c, dtype=np.float64)
    return DifferentialExpr([(REACTION, 1.0, c.reshape(-1) if c.ndim else c.reshape(1), c.ndim >= 1 and c.size > 1)])


# ---- geometry / space -----------------------------------------------------------------------------------------------
class Triangulation:
    """Triangulation<M,N>: nodes n_nodes x N, cells n_cells x (M+1), boundary markers per node."""

    def __init__(self, nodes, cells, boundary):
        self.nodes = np.asarray(nodes, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.boundary = np.ascontiguousarray(boundary, dtype=np.uint8).ravel()
        self.local_dim = self.cells.shape[1] - 1
        self.embed_dim = self.nodes.shape[1]

    def n_nodes(self):
        return self.nodes.shape[0]

    def n_cells(self):
        return self.cells.shape[0]


def mesh_topology(mesh):
    """The arrays the reference's Triangulation constructors build (triangulation.h:143-196, 319-399), computed on the
    device (fdb_topology_*).  Same dict layout as oracle.mesh_topology."""
    M = mesh.local_dim
    n_cells, nv = mesh.n_cells(), M + 1
    h = C.c_void_p()
    _check(lib().fdb_topology_create(C.byref(h), M, mesh.n_nodes(), n_cells, _ptr(mesh.cells), _ptr(mesh.boundary)))
    try:
        nf, ne, nec = C.c_int(), C.c_int(), C.c_int64()
        _check(lib().fdb_topology_sizes(h, C.byref(nf), C.byref(ne), C.byref(nec)))
        nf, ne, nec = nf.value, ne.value, nec.value
        out = {"neighbors": np.empty((n_cells, nv), np.int32), "facets": np.empty((nf, M), np.int32),
               "cell_to_facets": np.empty((n_cells, nv), np.int32), "facet_to_cells": np.empty((nf, 2), np.int32),
               "facet_boundary": np.empty(nf, np.uint8), "n_facets": nf, "n_edges": ne}
        if M == 3:
            out.update(edges=np.empty((ne, 2), np.int32), face_to_edges=np.empty((nf, 3), np.int32),
                       edge_boundary=np.empty(ne, np.uint8), edge_cell_ptr=np.empty(ne + 1, np.int32),
                       edge_cells=np.empty(nec, np.int32))
        g = lambda k: _ptr(out[k]) if k in out else None
        _check(lib().fdb_topology_download(h, g("neighbors"), g("facets"), g("cell_to_facets"), g("facet_to_cells"),
                                           g("facet_boundary"), g("edges"), g("face_to_edges"), g("edge_boundary"),
                                           g("edge_cell_ptr"), g("edge_cells")))
    finally:
        lib().fdb_topology_destroy(h)
    return out


class LagrangianBasis:
    """LagrangianBasis<Mesh,R>: DOF table built on the device (fdb_enumerate_dofs)."""

    def __init__(self, mesh, order):
        self.mesh, self.order = mesh, order
        M = mesh.local_dim
        nv = M + 1
        nb = nv if order == 1 else nv * (nv + 1) // 2
        n_cells = mesh.n_cells()
        dofs = np.zeros((n_cells, nb), dtype=np.int32, order="F")
        cap = mesh.n_nodes() + n_cells * (3 if M == 2 else 6)
        bd = np.zeros(cap, dtype=np.uint8)
        n = C.c_int()
        _check(lib().fdb_enumerate_dofs(M, order, mesh.n_nodes(), n_cells, _ptr(mesh.cells), _ptr(mesh.boundary),
                                        _ptr(dofs), _ptr(bd), C.byref(n)))
        self._dofs, self._size, self._boundary = dofs, n.value, bd[:n.value].copy()

    def size(self):
        return self._size

    def dofs(self):
        return self._dofs

    def boundary_dofs(self):
        return self._boundary


class Vector:
    def __init__(self, n, host=None):
        self.n = int(n)
        self.h = C.c_void_p()
        _check(lib().fdb_vector_create(C.c_int64(self.n), C.byref(self.h)))
        if host is not None:
            self.upload(host)

    def upload(self, host):
        a = np.ascontiguousarray(host, dtype=np.float64).ravel()
        _check(lib().fdb_vector_upload(self.h, _ptr(a), C.c_int64(a.size)))
        return self

    def download(self):
        out = np.empty(self.n)
        _check(lib().fdb_vector_download(self.h, _ptr(out), C.c_int64(self.n)))
        return out

    def fill(self, v):
        _check(lib().fdb_vector_fill(self.h, C.c_double(v)))
        return self

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.fdb_vector_destroy(self.h)
            self.h = None


class Space:
    """Device-resident mesh + FE space (fdb_space)."""

    def __init__(self, mesh, order, dofs, n_dofs, boundary_dofs=None, pass_cells=False):
        self.mesh, self.order, self.n_dofs = mesh, order, int(n_dofs)
        M, N = mesh.local_dim, mesh.embed_dim
        nodes_cm = np.asfortranarray(mesh.nodes)
        dofs_cm = np.asfortranarray(np.asarray(dofs, dtype=np.int32))
        self.h = C.c_void_p()
        _check(lib().fdb_space_create(C.byref(self.h), M, N, order, mesh.n_nodes(), mesh.n_cells(), _ptr(nodes_cm),
                                      _ptr(mesh.cells) if pass_cells else None, self.n_dofs, _ptr(dofs_cm)))
        nb, nq = C.c_int(), C.c_int()
        _check(lib().fdb_space_info(self.h, None, None, C.byref(nb), C.byref(nq)))
        self.n_basis, self.n_quad = nb.value, nq.value
        if boundary_dofs is not None:
            self.set_boundary(boundary_dofs)

    def set_boundary(self, boundary_dofs):
        b = np.ascontiguousarray(boundary_dofs, dtype=np.uint8).ravel()
        assert b.size == self.n_dofs
        _check(lib().fdb_space_set_boundary(self.h, _ptr(b)))

    def set_dof0_rule(self, on):
        _check(lib().fdb_space_set_dof0_rule(self.h, int(on)))

    def set_stream(self, cuda_stream_ptr):
        _check(lib().fdb_space_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_fused(self, on=True):
        _check(lib().fdb_space_set_fused(self.h, int(on)))

    def set_profiling(self, on=True):
        _check(lib().fdb_space_set_profiling(self.h, int(on)))

    def last_timings(self):
        ms = (C.c_double * 4)()
        n = C.c_int()
        _check(lib().fdb_space_last_timings(self.h, ms, 4, C.byref(n)))
        return [ms[k] for k in range(n.value)]

    def last_path(self):
        """(fused, launches) of the last assembly on this space."""
        f, n = C.c_int(), C.c_int()
        _check(lib().fdb_space_last_path(self.h, C.byref(f), C.byref(n)))
        return bool(f.value), n.value

    def last_kernel(self):
        """2: persistent fused kernel, 1: fused kernel, 0: contribution list + segmented reduction (last assembly)."""
        f, n = C.c_int(), C.c_int()
        _check(lib().fdb_space_last_path(self.h, C.byref(f), C.byref(n)))
        return f.value

    def sync(self):
        _check(lib().fdb_space_sync(self.h))

    def prepare(self, symmetric=True):
        _check(lib().fdb_space_prepare(self.h, int(symmetric)))

    def prepare_pattern(self, symmetric=True):
        """Builds the sparsity pattern + scatter map of one symmetry class (no fused plan); returns nnz."""
        nnz = C.c_int64()
        _check(lib().fdb_pattern_nnz(self.h, int(symmetric), C.byref(nnz)))
        return nnz.value

    def pattern(self, symmetric):
        nnz = C.c_int64()
        _check(lib().fdb_pattern_nnz(self.h, int(symmetric), C.byref(nnz)))
        outer = np.empty(self.n_dofs + 1, dtype=np.int32)
        inner = np.empty(nnz.value, dtype=np.int32)
        _check(lib().fdb_pattern_download(self.h, int(symmetric), _ptr(outer), _ptr(inner)))
        return outer, inner

    def quadrature_nodes(self):
        out = np.empty((self.mesh.n_cells() * self.n_quad, self.mesh.embed_dim), order="F")
        _check(lib().fdb_quadrature_nodes(self.h, _ptr(out)))
        return out

    def dofs_coords(self):
        out = np.empty((self.n_dofs, self.mesh.embed_dim), order="F")
        _check(lib().fdb_dofs_coords(self.h, _ptr(out)))
        return out

    # ---- next-row N1: point location and basis evaluation (lagrangian_basis.h:203-283) -------------------------------
    def locate(self, locs):
        """Triangulation::locate: id of a cell containing each point (smallest id if shared), -1 outside the domain."""
        L = np.asfortranarray(np.asarray(locs, dtype=np.float64))
        assert L.ndim == 2 and L.shape[1] == self.mesh.embed_dim
        ids = np.empty(L.shape[0], dtype=np.int32)
        _check(lib().fdb_locate(self.h, C.c_int64(L.shape[0]), _ptr(L), _ptr(ids)))
        return ids

    def eval_pointwise(self, locs):
        """pointwise_evaluation::eval: (cell ids, cols n_locs x n_basis, vals n_locs x n_basis), the triplets of Psi in
        the reference's emission order (cols == -1: point outside the domain)."""
        L = np.asfortranarray(np.asarray(locs, dtype=np.float64))
        assert L.ndim == 2 and L.shape[1] == self.mesh.embed_dim
        n = L.shape[0]
        ids = np.empty(n, dtype=np.int32)
        cols = np.empty((n, self.n_basis), dtype=np.int32)
        vals = np.empty((n, self.n_basis))
        _check(lib().fdb_eval_pointwise(self.h, C.c_int64(n), _ptr(L), _ptr(ids), _ptr(cols), _ptr(vals)))
        return ids, cols, vals

    def eval_areal(self, incidence):
        """areal_evaluation::eval: (rows, cols, vals, D) -- triplets in emission order (duplicates unsummed) and the
        subdomain measures."""
        inc = np.asfortranarray(np.asarray(incidence, dtype=np.float64))
        assert inc.ndim == 2 and inc.shape[1] == self.mesh.n_cells()
        nt = C.c_int64()
        _check(lib().fdb_eval_areal(self.h, inc.shape[0], _ptr(inc), C.c_int64(0), C.byref(nt), None, None, None, None))
        n = max(nt.value, 1)
        rows, cols, vals = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32), np.empty(n)
        D = np.empty(inc.shape[0])
        _check(lib().fdb_eval_areal(self.h, inc.shape[0], _ptr(inc), C.c_int64(n), C.byref(nt), _ptr(rows), _ptr(cols),
                                    _ptr(vals), _ptr(D)))
        return rows[:nt.value], cols[:nt.value], vals[:nt.value], D

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.fdb_space_destroy(self.h)
            self.h = None


class SolverOptions:
    def __init__(self, kind="cg", rtol=1e-8, maxit=0, jacobi=False, check_every=0):
        self.kind, self.rtol, self.maxit, self.jacobi, self.check_every = kind, rtol, maxit, jacobi, check_every

    def c(self):
        o = _SolverOpts()
        o.kind = 0 if self.kind == "cg" else 1
        o.jacobi, o.maxit, o.check_every, o.rtol = int(self.jacobi), int(self.maxit), int(self.check_every), self.rtol
        return o


class Matrix:
    """An assembled operator on the device (fdb_matrix)."""

    def __init__(self, space):
        self.space = space
        self.h = C.c_void_p()
        _check(lib().fdb_matrix_create(space.h, C.byref(self.h)))

    def assemble(self, op, symmetric=None):
        d = op.descriptor(self.space.mesh.n_cells() * self.space.n_quad, symmetric)
        _check(lib().fdb_assemble_operator(self.space.h, C.byref(d), self.h))
        return self

    def nnz(self):
        n = C.c_int64()
        _check(lib().fdb_matrix_nnz(self.h, C.byref(n)))
        return n.value

    def download_csc(self, pinned=False):
        nnz = self.nnz()
        empty = pinned_empty if pinned else np.empty
        outer = empty(self.space.n_dofs + 1, dtype=np.int32)
        inner = empty(nnz, dtype=np.int32)
        val = empty(nnz, dtype=np.float64)
        _check(lib().fdb_matrix_download_csc(self.h, _ptr(outer), _ptr(inner), _ptr(val)))
        return outer, inner, val

    def set_dirichlet(self, g, b, x0=None):
        _check(lib().fdb_set_dirichlet(self.h, g.h, b.h, x0.h if x0 is not None else None))

    def set_partition(self, comm, local):
        """local: partition.LocalProblem of this rank"""
        nb = np.ascontiguousarray(local.neighbors, dtype=np.int32)
        sc = np.ascontiguousarray(local.send_counts, dtype=np.int32)
        si = np.ascontiguousarray(local.send_idx, dtype=np.int32)
        rc = np.ascontiguousarray(local.recv_counts, dtype=np.int32)
        self._comm = comm
        _check(lib().fdb_matrix_set_partition(self.h, comm.h, int(local.n_owned), int(nb.size), _ptr(nb), _ptr(sc),
                                              _ptr(si), _ptr(rc)))

    def enable_peer_memory(self, local, all_gather):
        """Sets up the peer-memory plan of the persistent multi-GPU CG.  `all_gather(obj)` returns the list of every
        rank's obj (e.g. torch.distributed.all_gather_object)."""
        h = C.create_string_buffer(64)
        _check(lib().fdb_matrix_peer_export(self.h, h))
        nb = [int(q) for q in local.neighbors]
        recv_off = np.concatenate([[0], np.cumsum(local.recv_counts)]).astype(np.int64)
        info = all_gather({"handle": bytes(h.raw), "n_halo": int(recv_off[-1]), "nbr": nb,
                           "recv_off": [int(v) for v in recv_off]})
        handles = b"".join(i["handle"] for i in info)
        n_halo = np.array([i["n_halo"] for i in info], dtype=np.int64)
        off = np.array([info[q]["recv_off"][info[q]["nbr"].index(local.rank)] for q in nb], dtype=np.int32)
        _check(lib().fdb_matrix_peer_connect(self.h, C.c_char_p(handles), _ptr(n_halo), _ptr(off)))

    def spmv(self, x, y):
        _check(lib().fdb_spmv(self.h, x.h, y.h))

    def solve(self, b, x, opts, raise_on_fail=True):
        st = _SolveStats()
        o = opts.c()
        rc = lib().fdb_solve(self.h, b.h, x.h, C.byref(o), C.byref(st))
        if rc != FDB_OK and (raise_on_fail or rc != FDB_ERR_NOT_CONVERGED):
            _check(rc)
        return {"iters": st.iters, "converged": bool(st.converged), "rel_resid": st.rel_resid, "seconds": st.seconds}

    def solve_host(self, b, x0, opts, raise_on_fail=True):
        st = _SolveStats()
        o = opts.c()
        bb = np.ascontiguousarray(b, dtype=np.float64).ravel()
        x = np.ascontiguousarray(x0, dtype=np.float64).ravel().copy()
        rc = lib().fdb_solve_host(self.h, _ptr(bb), _ptr(x), C.byref(o), C.byref(st))
        if rc != FDB_OK and (raise_on_fail or rc != FDB_ERR_NOT_CONVERGED):
            _check(rc)
        return x, {"iters": st.iters, "converged": bool(st.converged), "rel_resid": st.rel_resid,
                   "seconds": st.seconds}

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.fdb_matrix_destroy(self.h)
            self.h = None


def solve_parabolic(stiff, mass, dt, f_quad, g, u0, opts):
    """FEMLinearParabolicSolver::solve (fem_linear_parabolic_solver.h:37-72) on the device.
    f_quad: (n_cells*nq) x m, g: n_dofs x m or None, u0: n_dofs.  Returns (solution n_dofs x m, stats)."""
    f = np.asfortranarray(f_quad, dtype=np.float64)
    m = f.shape[1]
    n = stiff.space.n_dofs
    gg = None if g is None else np.asfortranarray(g, dtype=np.float64)
    u = np.ascontiguousarray(u0, dtype=np.float64).ravel()
    sol = np.zeros((n, m), order="F")
    st = _SolveStats()
    o = opts.c()
    rc = lib().fdb_solve_parabolic(stiff.h, mass.h, C.c_double(dt), m, _ptr(f), _ptr(gg), _ptr(u), _ptr(sol), C.byref(o),
                                   C.byref(st))
    if rc not in (FDB_OK, FDB_ERR_NOT_CONVERGED):
        _check(rc)
    return sol, {"iters": st.iters, "converged": bool(st.converged), "rel_resid": st.rel_resid, "seconds": st.seconds}


class Comm:
    """NCCL communicator of the multi-GPU solve (fdb_comm).  `broadcast(obj)` is any host-side broadcast from rank 0
    (e.g. torch.distributed.broadcast_object_list) used once to ship the 128-byte NCCL id."""

    def __init__(self, rank, world, broadcast):
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            _check(lib().fdb_comm_unique_id(idbuf))
        raw = broadcast(bytes(idbuf.raw))
        self.rank, self.world = rank, world
        self.h = C.c_void_p()
        _check(lib().fdb_comm_create(C.byref(self.h), rank, world, C.c_char_p(raw)))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.fdb_comm_destroy(self.h)
            self.h = None


class Assembler:
    """Assembler<FEM, D, B, I>(mesh, integrator, n_dofs, dofs) -- fem_assembler.h:46-49.  Host arrays in, host arrays
    out, exactly like the reference (the integrator is implied by (M, R): integrator_tables.h:23-58)."""

    def __init__(self, mesh, order, n_dofs, dofs):
        self.space = Space(mesh, order, dofs, n_dofs)
        self._pattern = {}   # symmetry class -> (outer, inner): every operator of a class shares one pattern

    def discretize_operator(self, op, symmetric=None):
        """Returns the CSC arrays (outer, inner, values) of the SpMatrix<double> the reference returns.  The arrays
        live in page-locked memory (DMA targets).  The index arrays of a symmetry class are downloaded once per
        Assembler: a second operator on the same space (FEMSolverBase::init assembles stiff and mass,
        fem_solver_base.h:113,136) only moves its values."""
        s = self.space
        sym = bool(op.is_symmetric if symmetric is None else symmetric)
        nnz = C.c_int64()
        _check(lib().fdb_pattern_nnz(s.h, int(sym), C.byref(nnz)))
        val = pinned_empty(nnz.value, np.float64)
        d = op.descriptor(s.mesh.n_cells() * s.n_quad, sym)
        if sym in self._pattern:
            outer, inner = self._pattern[sym]
            _check(lib().fdb_discretize_operator(s.h, C.byref(d), None, None, _ptr(val)))
        else:
            outer = pinned_empty(s.n_dofs + 1, np.int32)
            inner = pinned_empty(nnz.value, np.int32)
            _check(lib().fdb_discretize_operator(s.h, C.byref(d), _ptr(outer), _ptr(inner), _ptr(val)))
            self._pattern[sym] = (outer, inner)
        return outer, inner, val

    def discretize_forcing(self, f_quad):
        s = self.space
        f = np.ascontiguousarray(f_quad, dtype=np.float64).ravel()
        assert f.size == s.mesh.n_cells() * s.n_quad
        b = np.empty(s.n_dofs)
        _check(lib().fdb_discretize_forcing(s.h, _ptr(f), _ptr(b)))
        return b


class PDE:
    """PDE<D,E,F,FEM,fem_order<R>> facade (pde/pde.h:40-114) over FEMSolverBase::init / set_dirichlet_bc /
    FEMLinearEllipticSolver::solve (fem_solver_base.h:106-155, fem_linear_elliptic_solver.h:34-50)."""

    def __init__(self, mesh, L, order=1, forcing=None, solver=None):
        self.mesh, self.L, self.order = mesh, L, order
        self.basis = LagrangianBasis(mesh, order)
        self.space = Space(mesh, order, self.basis.dofs(), self.basis.size(), self.basis.boundary_dofs())
        self.forcing, self.bc = forcing, None
        self.solver = solver or SolverOptions("cg" if L.is_symmetric else "bicgstab")
        self.is_init, self.success = False, False

    def n_dofs(self):
        return self.space.n_dofs

    def dof_coords(self):
        return self.space.dofs_coords()

    def quadrature_nodes(self):
        return self.space.quadrature_nodes()

    def set_forcing(self, f):
        self.forcing = f

    def set_dirichlet_bc(self, g):
        self.bc = np.ascontiguousarray(g, dtype=np.float64).ravel()

    def init(self):  # fem_solver_base.h:106-139: stiff, force, mass
        s = self.space
        self._stiff = Matrix(s).assemble(self.L)
        f = self.forcing
        if callable(f):
            f = f(self.quadrature_nodes())
        self._force = Vector(s.n_dofs)
        fq = Vector(s.mesh.n_cells() * s.n_quad, f)
        _check(lib().fdb_assemble_forcing(s.h, fq.h, self._force.h))
        self._mass = Matrix(s).assemble(reaction(1.0))
        self.is_init = True

    def solve(self):  # pde.h:102-105
        if not self.is_init:
            raise RuntimeError("solver must be initialized first!")
        s = self.space
        x = Vector(s.n_dofs).fill(0.0)
        if self.bc is not None:
            g = Vector(s.n_dofs, self.bc)
            self._stiff.set_dirichlet(g, self._force, x)
        st = self._stiff.solve(self._force, x, self.solver, raise_on_fail=False)
        self.success = st["converged"]
        self.stats = st
        self._solution = x.download()

    def solution(self):
        return self._solution

    def stiff(self):
        return self._stiff.download_csc()

    def mass(self):
        return self._mass.download_csc()

    def force(self):
        return self._force.download()
