"""ctypes front-end of the CPU oracle (oracle/fdapde_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under fdapde-core_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfdapde_oracle.so")

LAPLACIAN, DIFFUSION, ADVECTION, REACTION, DT = 0, 1, 2, 3, 4


class _Term(C.Structure):
    _fields_ = [("kind", C.c_int32), ("space_varying", C.c_int32), ("scale", C.c_double),
                ("coeff", C.c_void_p)]


def build(force=False):
    src = os.path.join(_HERE, "fdapde_oracle.c")
    so_mt = os.path.join(_HERE, "_build", "libfdapde_oracle_mt.so")
    if force or any(not os.path.exists(f) or os.path.getmtime(f) < os.path.getmtime(src) for f in (_SO, so_mt)):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_poly_eval.restype = C.c_double
        L.orc_poly_grad.restype = C.c_double
        L.orc_integrate_one.restype = C.c_double
        L.orc_assemble_operator.restype = C.c_int64
        L.orc_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def n_basis(M, R):
    return lib().orc_n_basis(M, R)


def n_quad(M, R):
    return lib().orc_n_quad(M, R)


def quadrature(M, R):
    nq = n_quad(M, R)
    nodes = np.zeros((nq, M))
    w = np.zeros(nq)
    assert lib().orc_quadrature(M, R, _p(nodes), _p(w)) == 0
    return nodes, w


def quadrature_table(M, K):
    nodes = np.zeros((K, M))
    w = np.zeros(K)
    assert lib().orc_quadrature_table(M, K, _p(nodes), _p(w)) == 0
    return nodes, w


def ref_basis_coeffs(M, R):
    nb = n_basis(M, R)
    c = np.zeros((nb, nb))
    lib().orc_ref_basis_coeffs(M, R, _p(c))
    return c


def poly_table(M, R):
    nb = n_basis(M, R)
    t = np.zeros((nb, M), dtype=np.int32)
    lib().orc_poly_table(M, R, _p(t))
    return t


def reference_nodes(M, R):
    nb = n_basis(M, R)
    t = np.zeros((nb, M))
    assert lib().orc_reference_nodes(M, R, _p(t)) == 0
    return t


def poly_eval(M, R, coeff, p):
    coeff, p = _f64(coeff), _f64(p)
    return lib().orc_poly_eval(M, R, _p(coeff), _p(p))


def poly_grad(M, R, coeff, d, p):
    coeff, p = _f64(coeff), _f64(p)
    return lib().orc_poly_grad(M, R, _p(coeff), d, _p(p))


def cell_geometry(v):
    """v: (M+1) x N vertex coordinates. Returns J (N x M), invJ (M x N), measure."""
    v = _f64(v)
    M, N = v.shape[0] - 1, v.shape[1]
    J = np.zeros((N, M))
    invJ = np.zeros((M, N))
    meas = C.c_double()
    assert lib().orc_cell_geometry(M, N, _p(v), _p(J), _p(invJ), C.byref(meas)) == 0
    return J, invJ, meas.value


class Terms:
    """Flattened operator expression: list of (kind, scale, coeff, space_varying)."""

    def __init__(self, terms):
        self.keep = []
        self.n = len(terms)
        self.arr = (_Term * max(self.n, 1))()
        for k, t in enumerate(terms):
            kind, scale = t[0], t[1]
            coeff = t[2] if len(t) > 2 else None
            sv = int(t[3]) if len(t) > 3 else 0
            self.arr[k].kind = kind
            self.arr[k].scale = scale
            self.arr[k].space_varying = sv
            if coeff is not None:
                a = _f64(np.asarray(coeff, dtype=np.float64))
                if kind == DIFFUSION and not sv:
                    a = _f64(np.asarray(coeff, dtype=np.float64).T.ravel())  # column-major K
                self.keep.append(a)
                self.arr[k].coeff = a.ctypes.data
            else:
                self.arr[k].coeff = None


def local_matrix(M, R, v, terms, cell_id=0):
    v = _f64(v)
    N = v.shape[1]
    nb = n_basis(M, R)
    out = np.zeros((nb, nb))
    T = terms if isinstance(terms, Terms) else Terms(terms)
    assert lib().orc_local_matrix(M, N, R, _p(v), cell_id, T.n, T.arr, _p(out)) == 0
    return out


def physical_gradients(M, R, v, p):
    v, p = _f64(v), _f64(p)
    N = v.shape[1]
    nb = n_basis(M, R)
    g = np.zeros((nb, N))
    assert lib().orc_physical_gradients(M, N, R, _p(v), _p(p), _p(g)) == 0
    return g


def _mesh_args(nodes, cells):
    """nodes: n_nodes x N array (any order) -> column-major buffer; cells: n_cells x (M+1) row-major."""
    nodes = np.asarray(nodes, dtype=np.float64)
    nodes_cm = np.asfortranarray(nodes)
    cells = _i32(cells)
    return nodes_cm, cells, nodes.shape[0], nodes.shape[1], cells.shape[0], cells.shape[1] - 1


def assemble_operator(R, nodes, cells, dofs, n_dofs, terms, symmetric):
    """Returns CSC (outer, inner, values) exactly as the reference's discretize_operator would."""
    nodes_cm, cells, n_nodes, N, n_cells, M = _mesh_args(nodes, cells)
    dofs_cm = np.asfortranarray(np.asarray(dofs, dtype=np.int32))
    T = terms if isinstance(terms, Terms) else Terms(terms)
    o, i, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
    nnz = lib().orc_assemble_operator(M, N, R, n_nodes, n_cells, _p(nodes_cm), _p(cells), n_dofs, _p(dofs_cm),
                                      T.n, T.arr, int(bool(symmetric)), C.byref(o), C.byref(i), C.byref(v))
    assert nnz >= 0
    outer = np.ctypeslib.as_array(C.cast(o, C.POINTER(C.c_int32)), (n_dofs + 1,)).copy()
    inner = np.ctypeslib.as_array(C.cast(i, C.POINTER(C.c_int32)), (max(nnz, 1),))[:nnz].copy()
    val = np.ctypeslib.as_array(C.cast(v, C.POINTER(C.c_double)), (max(nnz, 1),))[:nnz].copy()
    for ptr in (o, i, v):
        lib().orc_free(ptr)
    return outer, inner, val


_lib_mt = None


def lib_mt():
    """The all-core (OpenMP) build of the same source: only orc_assemble_operator_mt differs from the serial library."""
    global _lib_mt
    if _lib_mt is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "_build", "libfdapde_oracle_mt.so"))
        L.orc_assemble_operator_mt.restype = C.c_int64
        L.orc_free.argtypes = [C.c_void_p]
        _lib_mt = L
    return _lib_mt


def assemble_operator_mt(R, nodes, cells, dofs, n_dofs, terms, symmetric, n_threads=None):
    """Best-effort all-core CPU variant (bench.py --impl reference): bit-identical to assemble_operator."""
    nodes_cm, cells, n_nodes, N, n_cells, M = _mesh_args(nodes, cells)
    dofs_cm = np.asfortranarray(np.asarray(dofs, dtype=np.int32))
    T = terms if isinstance(terms, Terms) else Terms(terms)
    nt = int(n_threads or os.cpu_count() or 1)
    o, i, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
    L = lib_mt()
    nnz = L.orc_assemble_operator_mt(M, N, R, n_nodes, n_cells, _p(nodes_cm), _p(cells), n_dofs, _p(dofs_cm), T.n, T.arr,
                                     int(bool(symmetric)), nt, C.byref(o), C.byref(i), C.byref(v))
    assert nnz >= 0
    outer = np.ctypeslib.as_array(C.cast(o, C.POINTER(C.c_int32)), (n_dofs + 1,)).copy()
    inner = np.ctypeslib.as_array(C.cast(i, C.POINTER(C.c_int32)), (max(nnz, 1),))[:nnz].copy()
    val = np.ctypeslib.as_array(C.cast(v, C.POINTER(C.c_double)), (max(nnz, 1),))[:nnz].copy()
    for ptr in (o, i, v):
        L.orc_free(ptr)
    return outer, inner, val


def assemble_forcing(R, nodes, cells, dofs, n_dofs, f_quad):
    nodes_cm, cells, n_nodes, N, n_cells, M = _mesh_args(nodes, cells)
    dofs_cm = np.asfortranarray(np.asarray(dofs, dtype=np.int32))
    f = _f64(f_quad).ravel()
    assert f.size == n_cells * n_quad(M, R)
    b = np.zeros(n_dofs)
    assert lib().orc_assemble_forcing(M, N, R, n_nodes, n_cells, _p(nodes_cm), _p(cells), n_dofs, _p(dofs_cm),
                                      _p(f), _p(b)) == 0
    return b


def quadrature_nodes(R, nodes, cells):
    nodes_cm, cells, n_nodes, N, n_cells, M = _mesh_args(nodes, cells)
    out = np.zeros((n_cells * n_quad(M, R), N), order="F")
    assert lib().orc_quadrature_nodes(M, N, R, n_nodes, n_cells, _p(nodes_cm), _p(cells), _p(out)) == 0
    return out


def integrate_one(R, nodes, cells):
    nodes_cm, cells, n_nodes, N, n_cells, M = _mesh_args(nodes, cells)
    return lib().orc_integrate_one(M, N, R, n_nodes, n_cells, _p(nodes_cm), _p(cells))


def enumerate_edges(cells, boundary_nodes=None):
    cells = _i32(cells)
    n_cells, M = cells.shape[0], cells.shape[1] - 1
    ne = 3 if M == 2 else 6
    ce = np.zeros((n_cells, ne), dtype=np.int32)
    bn = None if boundary_nodes is None else np.ascontiguousarray(boundary_nodes, dtype=np.uint8)
    e, eb = C.c_void_p(), C.c_void_p()
    n_edges = lib().orc_enumerate_edges(M, n_cells, _p(cells), None if bn is None else _p(bn), _p(ce),
                                        C.byref(e), C.byref(eb))
    edges = np.ctypeslib.as_array(C.cast(e, C.POINTER(C.c_int32)), (max(n_edges, 1), 2))[:n_edges].copy()
    ebound = np.ctypeslib.as_array(C.cast(eb, C.POINTER(C.c_uint8)), (max(n_edges, 1),))[:n_edges].copy()
    lib().orc_free(e)
    lib().orc_free(eb)
    return ce, edges, ebound


def mesh_topology(cells, boundary_nodes=None):
    """Literal restatement of the Triangulation<2,N> / <3,3> constructors (triangulation.h:143-196, 319-399).
    Returns a dict: neighbors, facets (edges in 2D / faces in 3D), cell_to_facets, facet_to_cells, facet_boundary and,
    in 3D, edges, face_to_edges, edge_boundary, edge_cell_ptr, edge_cells."""
    cells = _i32(cells)
    n_cells, M = cells.shape[0], cells.shape[1] - 1
    nv, cap = M + 1, cells.shape[0] * (M + 1)
    bn = None if boundary_nodes is None else np.ascontiguousarray(boundary_nodes, dtype=np.uint8).ravel()
    nb = np.zeros((n_cells, nv), np.int32)
    facets = np.zeros((cap, M), np.int32)
    c2f = np.zeros((n_cells, nv), np.int32)
    f2c = np.zeros((cap, 2), np.int32)
    fb = np.zeros(cap, np.uint8)
    cap3 = 3 * cap if M == 3 else 1
    edges, f2e, eb = np.zeros((cap3, 2), np.int32), np.zeros((cap if M == 3 else 1, 3), np.int32), np.zeros(cap3, np.uint8)
    ecp, ec = np.zeros(cap3 + 1, np.int32), np.zeros(6 * n_cells if M == 3 else 1, np.int32)
    ne = C.c_int()
    nf = lib().orc_mesh_topology(M, n_cells, _p(cells), None if bn is None else _p(bn), _p(nb), _p(facets), _p(c2f),
                                 _p(f2c), _p(fb), _p(edges), _p(f2e), _p(eb), _p(ecp), _p(ec), C.byref(ne))
    out = {"neighbors": nb, "facets": facets[:nf].copy(), "cell_to_facets": c2f, "facet_to_cells": f2c[:nf].copy(),
           "facet_boundary": fb[:nf].copy(), "n_facets": nf, "n_edges": ne.value}
    if M == 3:
        out.update(edges=edges[:ne.value].copy(), face_to_edges=f2e[:nf].copy(), edge_boundary=eb[:ne.value].copy(),
                   edge_cell_ptr=ecp[:ne.value + 1].copy(), edge_cells=ec[:ecp[ne.value]].copy())
    return out


def enumerate_dofs(R, n_nodes, cells, boundary_nodes):
    """Returns (dofs n_cells x nb, n_dofs, boundary_dofs uint8[n_dofs])."""
    cells = _i32(cells)
    n_cells, M = cells.shape[0], cells.shape[1] - 1
    nb = n_basis(M, R)
    bn = np.ascontiguousarray(boundary_nodes, dtype=np.uint8).ravel()
    dofs = np.zeros((n_cells, nb), dtype=np.int32, order="F")
    max_edges = n_cells * (3 if M == 2 else 6)
    bd = np.zeros(n_nodes + max_edges, dtype=np.uint8)
    n_dofs = lib().orc_enumerate_dofs(M, R, n_nodes, n_cells, _p(cells), _p(bn), _p(dofs), _p(bd))
    return dofs, n_dofs, bd[:n_dofs].copy()


def dofs_coords(R, nodes, cells, dofs, n_dofs):
    nodes_cm, cells, n_nodes, N, n_cells, M = _mesh_args(nodes, cells)
    dofs_cm = np.asfortranarray(np.asarray(dofs, dtype=np.int32))
    out = np.zeros((n_dofs, N), order="F")
    assert lib().orc_dofs_coords(M, N, R, n_nodes, n_cells, _p(nodes_cm), _p(cells), n_dofs, _p(dofs_cm),
                                 _p(out)) == 0
    return out


def set_dirichlet(outer, inner, val, boundary_dofs, g, b):
    """In-place row replacement on CSC values and rhs (fem_solver_base.h:144-155)."""
    n = outer.size - 1
    bd = np.ascontiguousarray(boundary_dofs, dtype=np.uint8)
    g = _f64(g).ravel()
    assert val.flags.c_contiguous and b.flags.c_contiguous
    lib().orc_set_dirichlet(n, _p(_i32(outer)), _p(_i32(inner)), _p(val), _p(bd), _p(g), _p(b))


def cg(rowptr, colidx, val, b, x0, rtol=1e-8, maxit=100000, jacobi=False):
    n = rowptr.size - 1
    x = _f64(x0).copy()
    rel = C.c_double()
    it = lib().orc_cg(n, _p(_i32(rowptr)), _p(_i32(colidx)), _p(_f64(val)), _p(_f64(b)), _p(x), C.c_double(rtol),
                      maxit, int(jacobi), C.byref(rel))
    return x, it, rel.value


def bicgstab(rowptr, colidx, val, b, x0, rtol=1e-8, maxit=100000, jacobi=False):
    n = rowptr.size - 1
    x = _f64(x0).copy()
    rel = C.c_double()
    it = lib().orc_bicgstab(n, _p(_i32(rowptr)), _p(_i32(colidx)), _p(_f64(val)), _p(_f64(b)), _p(x),
                            C.c_double(rtol), maxit, int(jacobi), C.byref(rel))
    return x, it, rel.value


# ---- next-row N1: point location and basis evaluation matrices ----------------------------------------------------
def locate(nodes, cells, locs):
    nodes_cm, cells, n_nodes, N, n_cells, M = _mesh_args(nodes, cells)
    L = np.asfortranarray(np.asarray(locs, dtype=np.float64))
    ids = np.zeros(L.shape[0], dtype=np.int32)
    assert lib().orc_locate(M, N, n_nodes, n_cells, _p(nodes_cm), _p(cells), L.shape[0], _p(L), _p(ids)) == 0
    return ids


def eval_pointwise(R, nodes, cells, dofs, locs):
    """Returns (cell ids, cols n_locs x nb, vals n_locs x nb): the triplets of pointwise_evaluation in emission order
    (cols == -1 for points outside the domain)."""
    nodes_cm, cells, n_nodes, N, n_cells, M = _mesh_args(nodes, cells)
    dofs_cm = np.asfortranarray(np.asarray(dofs, dtype=np.int32))
    L = np.asfortranarray(np.asarray(locs, dtype=np.float64))
    nb = n_basis(M, R)
    ids = np.zeros(L.shape[0], dtype=np.int32)
    cols = np.zeros((L.shape[0], nb), dtype=np.int32)
    vals = np.zeros((L.shape[0], nb))
    assert lib().orc_eval_pointwise(M, N, R, n_nodes, n_cells, _p(nodes_cm), _p(cells), _p(dofs_cm), L.shape[0], _p(L),
                                    _p(ids), _p(cols), _p(vals)) == 0
    return ids, cols, vals


def eval_areal(R, nodes, cells, dofs, incidence):
    """Returns (rows, cols, vals, D): the triplets of areal_evaluation in emission order and the subdomain measures."""
    nodes_cm, cells, n_nodes, N, n_cells, M = _mesh_args(nodes, cells)
    dofs_cm = np.asfortranarray(np.asarray(dofs, dtype=np.int32))
    inc = np.asfortranarray(np.asarray(incidence, dtype=np.float64))
    assert inc.shape[1] == n_cells
    nb = n_basis(M, R)
    cap = int((inc == 1).sum()) * nb
    rows = np.zeros(max(cap, 1), dtype=np.int32)
    cols = np.zeros(max(cap, 1), dtype=np.int32)
    vals = np.zeros(max(cap, 1))
    D = np.zeros(inc.shape[0])
    lib().orc_eval_areal.restype = C.c_int64
    nt = lib().orc_eval_areal(M, N, R, n_nodes, n_cells, _p(nodes_cm), _p(cells), _p(dofs_cm), inc.shape[0], _p(inc),
                              _p(rows), _p(cols), _p(vals), _p(D))
    assert nt == cap
    return rows[:nt], cols[:nt], vals[:nt], D
