"""Row-block (element) partition of a mesh across ranks for the multi-GPU path.  Host side,
numpy only, no communication: every rank derives its own local mesh and the halo lists of BOTH directions from the
global mesh, and the lists agree between ranks because both sides sort them by global dof id.

Rank r owns the contiguous dof rows [bounds[r], bounds[r+1]).  Its local mesh is every cell that touches an owned row
(cells on the slab boundary are assembled by both neighbours, so assembly needs no communication and each owned row is
summed in the same ascending-cell order as on one GPU).  Local numbering = [owned dofs | halo dofs grouped by owner
rank, ascending global id inside a group].
"""
import numpy as np


def row_bounds(n_dofs, world):
    return np.array([n_dofs * r // world for r in range(world + 1)], dtype=np.int64)


class LocalProblem:
    """What one rank needs: local mesh arrays + the halo exchange plan for fdb_matrix_set_partition."""

    def __init__(self, rank, world, n_owned, own0, local_to_global, nodes, cells, boundary, cell_ids, neighbors,
                 send_counts, send_idx, recv_counts):
        self.rank, self.world = rank, world
        self.n_owned, self.own0 = n_owned, own0
        self.local_to_global = local_to_global
        self.nodes, self.cells, self.boundary = nodes, cells, boundary
        self.cell_ids = cell_ids
        self.neighbors, self.send_counts, self.send_idx, self.recv_counts = neighbors, send_counts, send_idx, recv_counts


def partition_p1(nodes, cells, boundary, rank, world, bounds=None):
    """P1 spaces (dofs == mesh nodes).  Returns the LocalProblem of `rank`."""
    n = nodes.shape[0]
    bounds = row_bounds(n, world) if bounds is None else np.asarray(bounds, dtype=np.int64)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    owner = (np.searchsorted(bounds, cells, side="right") - 1).astype(np.int32)  # owner rank of every cell vertex
    mine = owner == rank
    touch = mine.any(axis=1)
    cell_ids = np.nonzero(touch)[0]
    lc = cells[touch]
    lo = owner[touch]
    # halo dofs of this rank: non-owned vertices of its cells, grouped by owner
    halo = np.unique(lc[lo != rank])
    halo_owner = (np.searchsorted(bounds, halo, side="right") - 1).astype(np.int32)
    order = np.lexsort((halo, halo_owner))
    halo, halo_owner = halo[order], halo_owner[order]
    # dofs this rank must send to q: its owned vertices that share a cell with a vertex owned by q.
    # (any such cell touches a row of q, so it is in q's local mesh and the vertex is in q's halo)
    send_lists = {}
    mixed = touch & (~mine).any(axis=1)
    mc, mo = cells[mixed], owner[mixed]
    for q in np.unique(mo[mo != rank]):
        has_q = (mo == q).any(axis=1)
        v = mc[has_q][mo[has_q] == rank]
        send_lists[int(q)] = np.unique(v)
    neighbors = sorted(set(send_lists) | set(int(q) for q in np.unique(halo_owner)))
    send_counts, send_idx, recv_counts = [], [], []
    for q in neighbors:
        sl = send_lists.get(q, np.zeros(0, dtype=np.int64))
        send_counts.append(sl.size)
        send_idx.append(sl - r0)
        recv_counts.append(int((halo_owner == q).sum()))
    send_idx = np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, np.int32)
    local_to_global = np.concatenate([np.arange(r0, r1, dtype=np.int64), halo.astype(np.int64)])
    # global -> local for the vertices of the local cells
    g2l = np.full(n, -1, dtype=np.int64)
    g2l[local_to_global] = np.arange(local_to_global.size)
    local_cells = g2l[lc].astype(np.int32)
    assert (local_cells >= 0).all()
    lp = LocalProblem(rank, world, r1 - r0, r0, local_to_global, np.ascontiguousarray(nodes[local_to_global]),
                      local_cells, np.ascontiguousarray(boundary[local_to_global]), cell_ids,
                      np.array(neighbors, dtype=np.int32), np.array(send_counts, dtype=np.int32), send_idx,
                      np.array(recv_counts, dtype=np.int32))
    lp.dofs, lp.n_local_dofs, lp.owns_dof0, lp.node_ids = local_cells, int(local_to_global.size), r0 == 0, local_to_global
    return lp


def dof_owners(dofs, n_dofs, world, cell_bounds=None):
    """Owner rank of every dof for a general dof table (P2: vertex dofs first, edge dofs after, so contiguous id ranges
    are not geometric slabs -- SURVEY.md 8e "P2 caveat").  Cells are cut into `world` contiguous id ranges (slabs on
    structured meshes) and a dof belongs to the rank of its lowest-id incident cell."""
    n_cells = dofs.shape[0]
    cell_bounds = row_bounds(n_cells, world) if cell_bounds is None else np.asarray(cell_bounds, dtype=np.int64)
    first = np.full(n_dofs, n_cells, dtype=np.int64)
    ids = np.repeat(np.arange(n_cells, dtype=np.int64), dofs.shape[1])
    # repeated indices in a fancy assignment keep the last value written: feed the cells in descending order
    first[dofs.ravel()[::-1]] = ids[::-1]
    assert (first < n_cells).all(), "dof without an incident cell"
    return (np.searchsorted(cell_bounds, first, side="right") - 1).astype(np.int32)


def partition_dofs(nodes, cells, dofs, n_dofs, boundary_dofs, rank, world, owner=None):
    """General FE spaces (any dof table, e.g. P2).  Rank r owns the dofs `owner == r`; internally they are renumbered
    [owned dofs, ascending global id | halo dofs grouped by owner rank, ascending global id], so the device code sees the
    same contiguous-owned layout as for P1.  `local_to_global` maps back to the reference numbering (the ABI's vectors
    and matrices are always returned in it).  The local mesh lists every cell with an owned dof, in ascending global cell
    order, with its own compact node numbering."""
    dofs = np.asarray(dofs)
    owner = dof_owners(dofs, n_dofs, world) if owner is None else owner
    cell_owner = owner[dofs]                         # owner of every local dof slot
    mine = cell_owner == rank
    touch = mine.any(axis=1)
    cell_ids = np.nonzero(touch)[0]
    ld, lo = dofs[touch], cell_owner[touch]
    owned = np.nonzero(owner == rank)[0].astype(np.int64)
    halo = np.unique(ld[lo != rank]).astype(np.int64)
    halo_owner = owner[halo]
    order = np.lexsort((halo, halo_owner))
    halo, halo_owner = halo[order], halo_owner[order]
    local_to_global = np.concatenate([owned, halo])
    g2l = np.full(n_dofs, -1, dtype=np.int64)
    g2l[local_to_global] = np.arange(local_to_global.size)
    local_dofs = g2l[ld].astype(np.int32)
    assert (local_dofs >= 0).all()
    # send lists: owned dofs sharing a cell with a dof owned by q (that cell is in q's local mesh)
    send_lists = {}
    mixed = ~mine[touch].all(axis=1)
    mc, mo = ld[mixed], lo[mixed]
    for q in np.unique(mo[mo != rank]):
        has_q = (mo == q).any(axis=1)
        send_lists[int(q)] = np.unique(mc[has_q][mo[has_q] == rank])
    neighbors = sorted(set(send_lists) | set(int(q) for q in np.unique(halo_owner)))
    send_counts, send_idx, recv_counts = [], [], []
    for q in neighbors:
        sl = send_lists.get(q, np.zeros(0, dtype=np.int64))
        send_counts.append(sl.size)
        send_idx.append(g2l[sl])
        recv_counts.append(int((halo_owner == q).sum()))
    send_idx = np.concatenate(send_idx).astype(np.int32) if send_idx else np.zeros(0, np.int32)
    # compact node numbering of the local mesh (geometry only)
    lcells = cells[touch]
    used = np.unique(lcells)
    n2l = np.full(nodes.shape[0], -1, dtype=np.int64)
    n2l[used] = np.arange(used.size)
    lp = LocalProblem(rank, world, owned.size, int(owned[0]) if owned.size else -1, local_to_global,
                      np.ascontiguousarray(nodes[used]), n2l[lcells].astype(np.int32),
                      np.ascontiguousarray(np.asarray(boundary_dofs)[local_to_global]), cell_ids,
                      np.array(neighbors, dtype=np.int32), np.array(send_counts, dtype=np.int32), send_idx,
                      np.array(recv_counts, dtype=np.int32))
    lp.dofs = local_dofs
    lp.n_local_dofs = int(local_to_global.size)
    lp.owns_dof0 = bool(owned.size and owned[0] == 0)
    lp.node_ids = used
    return lp
