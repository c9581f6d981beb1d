"""Regenerates tests/golden/meshes.npz from the reference's own mesh fixtures.

Run in the dev container only (it reads /root/reference/test/data/mesh, which does not exist on the GPU box):
    python tests/golden/make_golden.py
File format restated from test/src/utils/mesh_loader.h:62-84 and fdaPDE/utils/IO/csv_reader.h:75-118:
header row, first column = row name, values possibly quoted, elements.csv 1-based.
"""
import os
import sys

import numpy as np

SRC = "/root/reference/test/data/mesh"
MESHES = ["c_shaped", "unit_square", "unit_square_16", "unit_square_32", "unit_sphere", "surface", "quasi_circle"]
MTX = "/root/reference/test/data/mtx"
PSI = ["lagrangian_pointwise_eval_order1", "lagrangian_pointwise_eval_order2", "lagrangian_areal_eval_order1",
       "lagrangian_areal_eval_order2"]


def read_mtx(path):
    """MatrixMarket coordinate real general -> (shape, rows, cols, vals), 0-based"""
    with open(path) as fh:
        lines = [ln for ln in fh if not ln.startswith("%")]
    nr, nc, nnz = (int(t) for t in lines[0].split())
    body = np.array([ln.split() for ln in lines[1:1 + nnz]])
    return (np.array([nr, nc]), body[:, 0].astype(np.int64) - 1, body[:, 1].astype(np.int64) - 1,
            body[:, 2].astype(np.float64))


def read_csv(path, dtype):
    rows = []
    with open(path) as fh:
        next(fh)
        for line in fh:
            parts = [p.strip().strip('"') for p in line.strip().split(",")][1:]
            rows.append([dtype(p) for p in parts])
    return np.array(rows, dtype=np.float64 if dtype is float else np.int32)


def main():
    out = {}
    for m in MESHES:
        d = os.path.join(SRC, m)
        out[m + "/points"] = read_csv(os.path.join(d, "points.csv"), float)
        out[m + "/elements"] = (read_csv(os.path.join(d, "elements.csv"), int) - 1).astype(np.int32)
        out[m + "/boundary"] = read_csv(os.path.join(d, "boundary.csv"), int).astype(np.uint8).ravel()
        out[m + "/n_edges_file"] = np.array(sum(1 for _ in open(os.path.join(d, "edges.csv"))) - 1)
        # topology fixtures (mesh_loader.h:73-79: 1-based, entries <= 0 mean "none"): neigh.csv = neighbours opposite to
        # each vertex, edges.csv = the facets (edges of triangles, faces of tetrahedra) in the generator's order
        ng = read_csv(os.path.join(d, "neigh.csv"), int)
        out[m + "/neigh"] = np.where(ng > 0, ng - 1, -1).astype(np.int32)
        out[m + "/facets_file"] = (read_csv(os.path.join(d, "edges.csv"), int) - 1).astype(np.int32)
        print(m, out[m + "/points"].shape, out[m + "/elements"].shape, int(out[m + "/boundary"].sum()),
              int(out[m + "/n_edges_file"]))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "meshes.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")
    # the synthetic generator of the package must reproduce the reference's structured meshes node for node (the CPU
    # convergence test and the benchmark ladders rely on it instead of shipping unit_square_64 / _128)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
    import __graft_entry__ as g
    fdb = g.load_package()
    for N in (16, 32, 64, 128):
        d = os.path.join(SRC, f"unit_square_{N}")
        n, c, b = fdb.meshes.unit_square(N)
        assert np.array_equal(n, read_csv(os.path.join(d, "points.csv"), float))
        assert np.array_equal(c, read_csv(os.path.join(d, "elements.csv"), int) - 1)
        assert np.array_equal(b.ravel(), read_csv(os.path.join(d, "boundary.csv"), int).ravel())
        print(f"unit_square_{N}: synthetic generator == reference files")
    # basis-evaluation fixtures (lagrangian_basis_test.cpp:200-238): locations, subdomain incidence, golden Psi
    psi = {"c_shaped/locs": read_csv(os.path.join(SRC, "c_shaped", "locs.csv"), float),
           "quasi_circle/incidence": read_csv(os.path.join(SRC, "quasi_circle", "incidence_matrix.csv"), float)}
    for name in PSI:
        shape, r, c, v = read_mtx(os.path.join(MTX, name + ".mtx"))
        psi[name + "/shape"], psi[name + "/rows"], psi[name + "/cols"], psi[name + "/vals"] = shape, r, c, v
        print(name, shape, v.size)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "psi.npz")
    np.savez_compressed(dst, **psi)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    sys.exit(main())
