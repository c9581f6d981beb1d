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
    return LocalProblem(rank, world, r1 - r0, r0, local_to_global, np.ascontiguousarray(nodes[local_to_global]),
                        local_cells, np.ascontiguousarray(boundary[local_to_global]), cell_ids,
                        np.array(neighbors, dtype=np.int32), np.array(send_counts, dtype=np.int32), send_idx,
                        np.array(recv_counts, dtype=np.int32))
