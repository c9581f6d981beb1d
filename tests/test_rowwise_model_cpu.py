"""Host-side model of the row-wise pattern build (csrc/pattern.cu: k_row_counts / k_row_fill), checked against the
definition it must reproduce: the emitted triplets of fem_assembler.h:79-110 sorted by (row, col, emission order), i.e.
what the sort build does on the device and what Eigen's setFromTriplets sums left to right (fem_assembler.h:112).
The CUDA kernels are compared with the sort build bit for bit on a GPU (tests/test_gpu_parity.py::
test_rowwise_pattern_build_equals_sort_build); this test pins the ALGORITHM (incidences sorted by (dof, cell), distinct
columns ranked per row, stored entries = head of the row for symmetric operators, stable counting placement, slot
numbering of the symmetric pair enumeration) without a device."""
import numpy as np
import pytest

from oracle import oracle as orc


def emitted_sorted(dofs, sym):
    """(pos, seg, ukeys) from the definition: every emitted triplet, sorted by (row, col, emission order)."""
    n_cells, nb = dofs.shape
    ne = nb * (nb + 1) // 2 if sym else nb * nb
    keys = []
    for e in range(n_cells):
        s = 0
        for i in range(nb):
            for j in range(i if sym else 0, nb):
                r, c = int(dofs[e, i]), int(dofs[e, j])
                if sym and r < c:
                    r, c = c, r
                keys.append((r, c, e * ne + s))
                s += 1
    keys.sort()
    pos, seg, uk = np.zeros(len(keys), np.int64), [], []
    for t, (r, c, i) in enumerate(keys):
        pos[i] = t
        if t == 0 or keys[t - 1][:2] != (r, c):
            seg.append(t)
            uk.append((r, c))
    return pos, np.array(seg + [len(keys)]), uk


def rowwise(dofs, n, sym):
    """The device algorithm, one row at a time."""
    n_cells, nb = dofs.shape
    ne = nb * (nb + 1) // 2 if sym else nb * nb
    key = dofs.ravel()                                  # emission order of the incidences: cell-major
    order = np.argsort(key, kind="stable")              # one stable sort by dof: cells ascend inside a dof
    inc = np.arange(n_cells * nb)[order]
    inc_ptr = np.searchsorted(key[order], np.arange(n + 1))
    pos, seg, uk, c0 = np.full(n_cells * ne, -1, np.int64), [], [], 0
    for r in range(n):
        cand = [(int(dofs[cid // nb, j]), cid // nb, cid % nb, j) for cid in inc[inc_ptr[r]:inc_ptr[r + 1]] for j in range(nb)]
        S = sorted({c[0] for c in cand})                # distinct columns, ranked
        nl = sum(1 for v in S if v <= r) if sym else len(S)   # stored entries = head of the full row
        cnt = [0] * nl
        for col, _, _, _ in cand:
            if not sym or col <= r:
                cnt[S.index(col)] += 1
        off = np.concatenate([[0], np.cumsum(cnt)]).astype(int)
        seg += [c0 + off[k] for k in range(nl)]
        uk += [(r, S[k]) for k in range(nl)]
        run = list(off[:-1])
        for col, cell, ai, j in cand:                   # emission order: stable counting placement
            if not sym or col <= r:
                k = S.index(col)
                if sym:
                    i, jj = min(ai, j), max(ai, j)
                    slot = i * nb - i * (i - 1) // 2 + (jj - i)
                else:
                    slot = ai * nb + j
                pos[cell * ne + slot] = c0 + run[k]
                run[k] += 1
        c0 += int(off[-1])
    return pos, np.array(seg + [c0]), uk


def _meshes():
    import __graft_entry__ as g
    return g.load_package().meshes


@pytest.mark.parametrize("case", ["cube_p1", "square_p1", "square_p2", "cube_p2"])
@pytest.mark.parametrize("sym", [True, False])
def test_rowwise_model_reproduces_the_sorted_triplets(case, sym):
    m = _meshes()
    nodes, cells, bnd = m.unit_cube(2) if case.startswith("cube") else m.unit_square(3)
    if case.endswith("p1"):
        dofs, n = np.asarray(cells), nodes.shape[0]
    else:
        dofs, n, _ = orc.enumerate_dofs(2, nodes.shape[0], cells, bnd)
        dofs = np.asarray(dofs)
    a, b = emitted_sorted(dofs, sym), rowwise(dofs, n, sym)
    assert np.array_equal(a[0], b[0]), "scatter map differs"
    assert np.array_equal(a[1], b[1]), "segment table differs"
    assert a[2] == b[2], "stored entries differ"
