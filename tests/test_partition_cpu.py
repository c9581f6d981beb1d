"""Host-side logic of the multi-GPU path, on CPU: the row-block partition, its halo plan, and a
world_size-2 gloo run proving that the two ranks' independently derived plans agree and that a distributed SpMV built
from them reproduces the global one."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _global_matrix(nodes, cells):
    from oracle import oracle as orc
    n = nodes.shape[0]
    o, i, v = orc.assemble_operator(1, nodes, cells, cells, n, [(orc.LAPLACIAN, -1.0), (orc.REACTION, 1.0, [1.0])], True)
    return sp.csc_matrix((v, i, o), shape=(n, n)).tocsr()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_partition_covers_rows_and_plans_agree(fdb, world):
    nodes, cells, bnd = fdb.meshes.unit_cube(6)
    n = nodes.shape[0]
    locs = [fdb.partition.partition_p1(nodes, cells, bnd, r, world) for r in range(world)]
    assert sum(l.n_owned for l in locs) == n
    for l in locs:
        # every cell touching an owned row is present, in ascending global cell order
        touch = ((cells >= l.own0) & (cells < l.own0 + l.n_owned)).any(axis=1)
        assert np.array_equal(l.cell_ids, np.nonzero(touch)[0])
        assert np.array_equal(l.local_to_global[l.cells], cells[l.cell_ids])
        assert np.array_equal(l.local_to_global[:l.n_owned], np.arange(l.own0, l.own0 + l.n_owned))
        # send list of r to q == halo segment of q owned by r, same order
        off = 0
        for k, q in enumerate(l.neighbors):
            sent = l.send_idx[off:off + l.send_counts[k]] + l.own0
            off += l.send_counts[k]
            lq = locs[q]
            kq = list(lq.neighbors).index(l.rank)
            roff = lq.n_owned + int(lq.recv_counts[:kq].sum())
            assert np.array_equal(sent, lq.local_to_global[roff:roff + lq.recv_counts[kq]])


def test_distributed_spmv_from_local_meshes_matches_global(fdb):
    from oracle import oracle as orc
    nodes, cells, bnd = fdb.meshes.unit_cube(5)
    n = nodes.shape[0]
    A = _global_matrix(nodes, cells)
    x = np.random.default_rng(0).standard_normal(n)
    y_ref = A @ x
    world = 4
    y = np.zeros(n)
    for r in range(world):
        l = fdb.partition.partition_p1(nodes, cells, bnd, r, world)
        nl = l.local_to_global.size
        o, i, v = orc.assemble_operator(1, l.nodes, l.cells, l.cells, nl, [(orc.LAPLACIAN, -1.0), (orc.REACTION, 1.0, [1.0])],
                                        True)
        Al = sp.csc_matrix((v, i, o), shape=(nl, nl)).tocsr()
        # owned rows of the local matrix are bit-identical to the global rows (same cells, same order)
        Ag = A[l.own0:l.own0 + l.n_owned][:, l.local_to_global]
        assert abs(Al[:l.n_owned] - Ag).max() == 0.0
        y[l.own0:l.own0 + l.n_owned] = (Al @ x[l.local_to_global])[:l.n_owned]
    assert np.max(np.abs(y - y_ref)) < 1e-13


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as g
    fdb = g.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nodes, cells, bnd = fdb.meshes.unit_cube(6)
    n = nodes.shape[0]
    l = fdb.partition.partition_p1(nodes, cells, bnd, rank, world)
    A = _global_matrix(nodes, cells)
    x = np.random.default_rng(1).standard_normal(n)
    # emulate the halo exchange with gloo point-to-point: send owned entries, receive the halo tail
    import torch
    xl = np.zeros(l.local_to_global.size)
    xl[:l.n_owned] = x[l.own0:l.own0 + l.n_owned]
    soff, roff, reqs, bufs = 0, l.n_owned, [], []
    for k, qk in enumerate(l.neighbors):
        sb = torch.from_numpy(xl[l.send_idx[soff:soff + l.send_counts[k]]].copy())
        rb = torch.zeros(int(l.recv_counts[k]), dtype=torch.float64)
        reqs += [dist.isend(sb, int(qk)), dist.irecv(rb, int(qk))]
        bufs.append((roff, rb))
        soff += l.send_counts[k]
        roff += int(l.recv_counts[k])
    for r_ in reqs:
        r_.wait()
    for off, rb in bufs:
        xl[off:off + rb.numel()] = rb.numpy()
    ok_halo = bool(np.array_equal(xl, x[l.local_to_global]))
    Al = A[l.own0:l.own0 + l.n_owned][:, l.local_to_global]
    y_own = Al @ xl
    # dot product = all-reduce of the owned partial sums
    t = torch.tensor([float(y_own @ xl[:l.n_owned])], dtype=torch.float64)
    dist.all_reduce(t)
    ok_dot = bool(abs(t.item() - float((A @ x) @ x)) < 1e-9 * abs(t.item()))
    ok_rows = bool(np.max(np.abs(y_own - (A @ x)[l.own0:l.own0 + l.n_owned])) < 1e-13)
    q.put((rank, ok_halo, ok_dot, ok_rows))
    dist.destroy_process_group()


def _gloo_worker_p2(rank, world, port, q):
    """Same exchange as _gloo_worker for a P2 space: owned dofs are not a contiguous id range, the plan is in the permuted
    (owned-first) numbering and results are mapped back through local_to_global."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    from oracle import oracle as orc
    fdb = g.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nodes, cells, bnd = fdb.meshes.unit_square(9)
    dofs, n_dofs, bd = orc.enumerate_dofs(2, nodes.shape[0], cells, bnd)
    terms = [(orc.LAPLACIAN, -1.0), (orc.ADVECTION, 1.0, [-1.0, 0.0]), (orc.REACTION, 1.0, [1.0])]   # C3's operator
    o, i, v = orc.assemble_operator(2, nodes, cells, dofs, n_dofs, terms, False)
    A = sp.csc_matrix((v, i, o), shape=(n_dofs, n_dofs)).tocsr()
    l = fdb.partition.partition_dofs(nodes, cells, dofs, n_dofs, bd, rank, world)
    ol, il, vl = orc.assemble_operator(2, l.nodes, l.cells, l.dofs, l.n_local_dofs, terms, False)
    Al = sp.csc_matrix((vl, il, ol), shape=(l.n_local_dofs, l.n_local_dofs)).tocsr()[:l.n_owned]
    x = np.random.default_rng(2).standard_normal(n_dofs)
    own_g = l.local_to_global[:l.n_owned]
    xl = np.zeros(l.n_local_dofs)
    xl[:l.n_owned] = x[own_g]
    soff, roff, reqs, bufs = 0, l.n_owned, [], []
    for k, qk in enumerate(l.neighbors):
        sb = torch.from_numpy(xl[l.send_idx[soff:soff + l.send_counts[k]]].copy())
        rb = torch.zeros(int(l.recv_counts[k]), dtype=torch.float64)
        reqs += [dist.isend(sb, int(qk)), dist.irecv(rb, int(qk))]
        bufs.append((roff, rb))
        soff += l.send_counts[k]
        roff += int(l.recv_counts[k])
    for r_ in reqs:
        r_.wait()
    for off, rb in bufs:
        xl[off:off + rb.numel()] = rb.numpy()
    ok_halo = bool(np.array_equal(xl, x[l.local_to_global]))
    y_own = Al @ xl
    ok_rows = bool(np.max(np.abs(y_own - (A @ x)[own_g])) < 1e-12)
    t = torch.tensor([float(y_own @ xl[:l.n_owned]), float(l.n_owned)], dtype=torch.float64)
    dist.all_reduce(t)
    ok_dot = bool(abs(t[0].item() - float((A @ x) @ x)) < 1e-9 * abs(t[0].item())) and int(t[1].item()) == n_dofs
    q.put((rank, ok_halo, ok_dot, ok_rows))
    dist.destroy_process_group()


def test_two_rank_gloo_p2_halo_exchange_and_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker_p2, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert r[1] and r[2] and r[3], r


def test_two_rank_gloo_halo_exchange_and_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert r[1] and r[2] and r[3], r


# ---- general dof tables (P2): owner = rank of the lowest-id incident cell, permuted owned-contiguous numbering ----------
def _p2_problem(fdb, dim):
    from oracle import oracle as orc
    if dim == 2:
        nodes, cells, bnd = fdb.meshes.unit_square(7)
    else:
        nodes, cells, bnd = fdb.meshes.unit_cube(3)
    dofs, n_dofs, bd = orc.enumerate_dofs(2, nodes.shape[0], cells, bnd)
    return nodes, cells, bnd, dofs, n_dofs, bd


@pytest.mark.parametrize("dim,world", [(2, 2), (2, 3), (3, 2), (3, 4)])
def test_p2_partition_plans_agree_and_rows_match(fdb, dim, world):
    from oracle import oracle as orc
    nodes, cells, bnd, dofs, n_dofs, bd = _p2_problem(fdb, dim)
    terms = [(orc.LAPLACIAN, -1.0), (orc.REACTION, 1.0, [1.0])]
    o, i, v = orc.assemble_operator(2, nodes, cells, dofs, n_dofs, terms, True)
    A = sp.csc_matrix((v, i, o), shape=(n_dofs, n_dofs)).tocsr()
    owner = fdb.partition.dof_owners(dofs, n_dofs, world)
    locs = [fdb.partition.partition_dofs(nodes, cells, dofs, n_dofs, bd, r, world, owner) for r in range(world)]
    assert sum(l.n_owned for l in locs) == n_dofs
    assert sum(l.owns_dof0 for l in locs) == 1
    x = np.random.default_rng(3).standard_normal(n_dofs)
    y = np.full(n_dofs, np.nan)
    for l in locs:
        own_g = l.local_to_global[:l.n_owned]
        assert np.array_equal(own_g, np.nonzero(owner == l.rank)[0])
        # local mesh: every cell with an owned dof, ascending global order, geometry and dof table consistent
        assert np.array_equal(l.cell_ids, np.nonzero((owner[dofs] == l.rank).any(axis=1))[0])
        assert np.array_equal(l.local_to_global[l.dofs], dofs[l.cell_ids])
        assert np.array_equal(l.nodes[l.cells], nodes[cells[l.cell_ids]])
        assert np.array_equal(l.boundary, np.asarray(bd)[l.local_to_global])
        # halo plan: what r sends to q is q's halo segment owned by r, in the same order
        off = 0
        for k, q in enumerate(l.neighbors):
            sent = l.local_to_global[l.send_idx[off:off + l.send_counts[k]]]
            assert (l.send_idx[off:off + l.send_counts[k]] < l.n_owned).all()
            off += l.send_counts[k]
            lq = locs[q]
            kq = list(lq.neighbors).index(l.rank)
            roff = lq.n_owned + int(lq.recv_counts[:kq].sum())
            assert np.array_equal(sent, lq.local_to_global[roff:roff + lq.recv_counts[kq]])
        # the matrix assembled from the local mesh reproduces the owned global rows (to rounding: the lower/upper
        # choice of a pair follows the local numbering) and a distributed SpMV reproduces the global one
        ol, il, vl = orc.assemble_operator(2, l.nodes, l.cells, l.dofs, l.n_local_dofs, terms, True)
        Al = sp.csc_matrix((vl, il, ol), shape=(l.n_local_dofs, l.n_local_dofs)).tocsr()
        Ag = A[own_g][:, l.local_to_global]
        assert abs(Al[:l.n_owned] - Ag).max() < 1e-14
        assert (Al[:l.n_owned] != 0).nnz == (Ag != 0).nnz
        y[own_g] = (Al @ x[l.local_to_global])[:l.n_owned]
    assert np.max(np.abs(y - A @ x)) < 1e-12


# ---- unstructured meshes: more than two neighbours per rank, irregular valence -------------------------------------------
def _delaunay_square(n_pts, seed):
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = np.concatenate([rng.random((n_pts, 2)), [[0, 0], [1, 0], [0, 1], [1, 1]]])
    tri = Delaunay(pts)
    cells = tri.simplices.astype(np.int32)
    # counter-clockwise or not does not matter to the path (|det J|); keep Delaunay's orientation
    from collections import Counter
    edges = Counter(tuple(sorted((c[a], c[b]))) for c in cells for a, b in ((0, 1), (0, 2), (1, 2)))
    bnd = np.zeros(pts.shape[0], dtype=np.uint8)
    for (a, b), k in edges.items():
        if k == 1:
            bnd[a] = bnd[b] = 1
    return pts, cells, bnd


@pytest.mark.parametrize("R,world,seed", [(1, 3, 0), (2, 3, 1), (2, 5, 2)])
def test_partition_on_unstructured_mesh(fdb, R, world, seed):
    """Random Delaunay triangulation with the cells numbered along a Z-order curve: ranks own compact patches and have
    several neighbours each; the halo plans must still agree pairwise and the distributed SpMV must match."""
    from oracle import oracle as orc
    pts, cells, bnd = _delaunay_square(300, seed)
    # cells along a Z-order curve: contiguous id ranges are compact 2D patches, not strips
    c = pts[cells].mean(axis=1)
    qx, qy = (c[:, 0] * 1023).astype(np.int64), (c[:, 1] * 1023).astype(np.int64)
    key = np.zeros(cells.shape[0], dtype=np.int64)
    for b in range(10):
        key |= ((qx >> b) & 1) << (2 * b) | ((qy >> b) & 1) << (2 * b + 1)
    cells = cells[np.argsort(key, kind="stable")]
    dofs, n_dofs, bd = orc.enumerate_dofs(R, pts.shape[0], cells, bnd)
    terms = [(orc.LAPLACIAN, -1.0), (orc.REACTION, 1.0, [1.0])]
    o, i, v = orc.assemble_operator(R, pts, cells, dofs, n_dofs, terms, True)
    A = sp.csc_matrix((v, i, o), shape=(n_dofs, n_dofs)).tocsr()
    owner = fdb.partition.dof_owners(dofs, n_dofs, world)
    locs = [fdb.partition.partition_dofs(pts, cells, dofs, n_dofs, bd, r, world, owner) for r in range(world)]
    assert sum(l.n_owned for l in locs) == n_dofs
    assert max(len(l.neighbors) for l in locs) >= 2
    x = np.random.default_rng(seed).standard_normal(n_dofs)
    y = np.full(n_dofs, np.nan)
    for l in locs:
        off = 0
        for k, q in enumerate(l.neighbors):
            sent = l.local_to_global[l.send_idx[off:off + l.send_counts[k]]]
            off += l.send_counts[k]
            lq = locs[q]
            kq = list(lq.neighbors).index(l.rank)                 # neighbour relation is symmetric
            roff = lq.n_owned + int(lq.recv_counts[:kq].sum())
            assert np.array_equal(sent, lq.local_to_global[roff:roff + lq.recv_counts[kq]])
        ol, il, vl = orc.assemble_operator(R, l.nodes, l.cells, l.dofs, l.n_local_dofs, terms, True)
        Al = sp.csc_matrix((vl, il, ol), shape=(l.n_local_dofs, l.n_local_dofs)).tocsr()
        own_g = l.local_to_global[:l.n_owned]
        assert abs(Al[:l.n_owned] - A[own_g][:, l.local_to_global]).max() < 1e-12
        y[own_g] = (Al @ x[l.local_to_global])[:l.n_owned]
    assert np.max(np.abs(y - A @ x)) < 1e-10
