"""Deterministic synthetic meshes of SURVEY.md section 8(d) (numpy, host side; inputs of the benchmark/tests)."""
import itertools

import numpy as np


def unit_square(N):
    """(N+1)^2 nodes, id = j*(N+1)+i -> (i/N, j/N); 2N^2 triangles (v00,v10,v11),(v00,v11,v01).
    Same ordering as the reference's test/data/mesh/unit_square_16 fixture."""
    i, j = np.meshgrid(np.arange(N + 1), np.arange(N + 1), indexing="xy")
    nodes = np.stack([i.ravel() / N, j.ravel() / N], axis=1).astype(np.float64)
    ci, cj = np.meshgrid(np.arange(N), np.arange(N), indexing="xy")
    v00 = (cj * (N + 1) + ci).ravel()
    t0 = np.stack([v00, v00 + 1, v00 + N + 2], axis=1)
    t1 = np.stack([v00, v00 + N + 2, v00 + N + 1], axis=1)
    cells = np.empty((2 * N * N, 3), dtype=np.int32)
    cells[0::2] = t0
    cells[1::2] = t1
    bi = (i.ravel() == 0) | (i.ravel() == N) | (j.ravel() == 0) | (j.ravel() == N)
    return nodes, cells, bi.astype(np.uint8)


def unit_cube(n):
    """(n+1)^3 nodes, id = (i(n+1)+j)(n+1)+k -> (i/n, j/n, k/n); every cube split into the 6 Kuhn tetrahedra
    (all monotone lattice paths from (i,j,k) to (i+1,j+1,k+1)); cube-major, then permutation order."""
    m = n + 1
    idx = np.arange(m)
    I, J, K = np.meshgrid(idx, idx, idx, indexing="ij")
    nodes = np.stack([I.ravel() / n, J.ravel() / n, K.ravel() / n], axis=1).astype(np.float64)
    ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    base = ((ci * m + cj) * m + ck).ravel().astype(np.int64)
    step = np.array([m * m, m, 1], dtype=np.int64)
    cells = np.empty((base.size, 6, 4), dtype=np.int32)
    for p, perm in enumerate(itertools.permutations(range(3))):
        v = base.copy()
        cells[:, p, 0] = v
        for s, axis in enumerate(perm):
            v = v + step[axis]
            cells[:, p, s + 1] = v
    cells = cells.reshape(-1, 4)
    b = (I.ravel() == 0) | (I.ravel() == n) | (J.ravel() == 0) | (J.ravel() == n) | (K.ravel() == 0) | (K.ravel() == n)
    return nodes, cells, b.astype(np.uint8)


def jitter(nodes, boundary, h, seed=0, amplitude=0.2):
    """Robustness variant: interior nodes moved by U(-amplitude*h, amplitude*h)."""
    rng = np.random.default_rng(seed)
    out = nodes.copy()
    interior = boundary == 0
    out[interior] += rng.uniform(-amplitude * h, amplitude * h, size=(int(interior.sum()), nodes.shape[1]))
    return out
