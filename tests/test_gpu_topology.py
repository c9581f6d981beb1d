"""Next-row N3: the topology arrays of the reference's Triangulation constructors (geometry/triangulation.h:143-196 for
triangles and surfaces, :319-399 for tetrahedra) built on the device by sort/unique (fdb_topology_*), bit-exact against
the literal hash-map restatement of the oracle and against the reference's neigh.csv / edges.csv fixtures."""
import os

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS2 = ["neighbors", "facets", "cell_to_facets", "facet_to_cells", "facet_boundary"]
KEYS3 = KEYS2 + ["edges", "face_to_edges", "edge_boundary", "edge_cell_ptr", "edge_cells"]


def check(fdb, nodes, cells, bnd):
    t = fdb.mesh_topology(fdb.Triangulation(nodes, cells, bnd))
    o = orc.mesh_topology(cells, bnd)
    assert t["n_facets"] == o["n_facets"] and t["n_edges"] == o["n_edges"]
    for k in (KEYS3 if cells.shape[1] == 4 else KEYS2):
        assert np.array_equal(t[k], o[k]), k
    return t


@pytest.mark.parametrize("mesh", ["c_shaped", "unit_square", "surface", "quasi_circle", "unit_sphere"])
def test_topology_matches_oracle_and_reference_fixtures(fdb, golden_meshes, mesh):
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))
    pts, els, bnd = golden_meshes(mesh)
    t = check(fdb, pts, els, bnd)
    assert np.array_equal(t["neighbors"], z[mesh + "/neigh"])                      # neigh.csv, exactly
    file_facets = np.sort(z[mesh + "/facets_file"][:, :els.shape[1] - 1], axis=1)
    assert set(map(tuple, t["facets"])) == set(map(tuple, file_facets))            # edges.csv, as a set


def test_topology_structured_meshes(fdb):
    for nodes, cells, bnd in (fdb.meshes.unit_square(33), fdb.meshes.unit_cube(7)):
        check(fdb, nodes, cells, bnd)
    # renumbered nodes (first-occurrence order no longer follows the node order)
    nodes, cells, bnd = fdb.meshes.unit_cube(5)
    perm = np.random.default_rng(11).permutation(nodes.shape[0])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    check(fdb, nodes[perm], inv[cells].astype(np.int32), np.asarray(bnd).ravel()[perm])


def test_topology_full_size_properties(fdb):
    """C4's mesh (10.1 M tetrahedra): counts from the Kuhn structure and consistency of the incidence arrays."""
    n = 119
    nodes, cells, bnd = fdb.meshes.unit_cube(n)
    t = fdb.mesh_topology(fdb.Triangulation(nodes, cells, bnd))
    nc = cells.shape[0]
    n_bfaces = 6 * 2 * n * n                                    # two triangles per boundary square
    assert t["n_facets"] == (4 * nc + n_bfaces) // 2 and int(t["facet_boundary"].sum()) == n_bfaces
    assert t["n_edges"] == 7 * n ** 3 + 9 * n ** 2 + 3 * n      # Kuhn mesh: n^3 diagonals + 3 n^2 (n+1) face diagonals + 3 n (n+1)^2 axis edges
    # Euler characteristic of the ball: V - E + F - C = 1
    assert nodes.shape[0] - t["n_edges"] + t["n_facets"] - nc == 1
    nb, c2f, f2c = t["neighbors"], t["cell_to_facets"], t["facet_to_cells"]
    assert int((nb < 0).sum()) == n_bfaces
    # neighbour relation is symmetric through the shared face
    i = np.arange(nc)
    for j in range(4):
        k = nb[:, j]
        ok = k >= 0
        shared = c2f[i[ok], 3 - j]                              # facet j' of a cell is opposite to vertex 3 - j'
        assert np.all((f2c[shared, 0] == i[ok]) | (f2c[shared, 1] == i[ok]))
        assert np.all((f2c[shared, 0] == k[ok]) | (f2c[shared, 1] == k[ok]))
    ptr = t["edge_cell_ptr"]
    assert ptr[-1] == 6 * nc and np.all(np.diff(ptr) >= 1)
