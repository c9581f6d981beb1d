"""Mesh and matrix file I/O in the reference's formats -- host side, numpy only.

CSV (fdaPDE/utils/IO/csv_reader.h:75-118): one header line; the first column of every row is a row name and is skipped;
tokens may be quoted (the reader drops every '"'); the reserved tokens NA / NaN / nan read as NaN.  A mesh directory
(test/src/utils/mesh_loader.h:62-84) holds points.csv, elements.csv (1-based), boundary.csv and optionally edges.csv /
neigh.csv (1-based, entries <= 0 mean "none").  MatrixMarket: coordinate real general, 1-based (Eigen::loadMarket)."""
import os

import numpy as np

_RESERVED = {"NA", "NaN", "nan"}


def read_csv(path, dtype=np.float64):
    rows = []
    with open(path) as fh:
        header = fh.readline()
        n_cols = len(header.rstrip("\n").split(",")) - 1
        for line in fh:
            line = line.rstrip("\n")
            if not line:
                continue
            tokens = [t.replace('"', "").strip() for t in line.split(",")][1:]
            rows.append([np.nan if t in _RESERVED else float(t) for t in tokens])
    out = np.array(rows, dtype=np.float64).reshape(len(rows), n_cols)
    if np.issubdtype(np.dtype(dtype), np.integer):
        return out.astype(dtype)
    return out


def write_csv(path, a, fmt="%.17g"):
    """Writes `a` the way R's write.csv laid out the reference fixtures: quoted header V1..Vn, quoted 1-based row names."""
    a = np.atleast_2d(np.asarray(a))
    integer = np.issubdtype(a.dtype, np.integer)
    with open(path, "w") as fh:
        fh.write(",".join(['""'] + [f'"V{j + 1}"' for j in range(a.shape[1])]) + "\n")
        for i, row in enumerate(a):
            vals = [str(int(v)) if integer else (fmt % v) for v in row]
            fh.write(",".join([f'"{i + 1}"'] + vals) + "\n")


def load_mesh(directory):
    """MeshLoader: returns dict(points, elements (0-based int32), boundary uint8[n_nodes], edges / neighbors when present
    (0-based, -1 = none))."""
    out = {"points": read_csv(os.path.join(directory, "points.csv")),
           "elements": (read_csv(os.path.join(directory, "elements.csv"), np.int64) - 1).astype(np.int32),
           "boundary": read_csv(os.path.join(directory, "boundary.csv"), np.int64).astype(np.uint8).ravel()}
    for name, key in (("edges.csv", "edges"), ("neigh.csv", "neighbors")):
        p = os.path.join(directory, name)
        if os.path.exists(p):
            a = read_csv(p, np.int64)
            out[key] = np.where(a > 0, a - 1, -1).astype(np.int32)
    return out


def save_mesh(directory, points, elements, boundary):
    os.makedirs(directory, exist_ok=True)
    write_csv(os.path.join(directory, "points.csv"), np.asarray(points, dtype=np.float64))
    write_csv(os.path.join(directory, "elements.csv"), np.asarray(elements, dtype=np.int64) + 1)
    write_csv(os.path.join(directory, "boundary.csv"), np.asarray(boundary, dtype=np.int64).reshape(-1, 1))


def read_mtx(path):
    """-> (shape, rows, cols, vals), 0-based, entries in file order"""
    with open(path) as fh:
        lines = [ln for ln in fh if ln.strip() and not ln.startswith("%")]
    nr, nc, nnz = (int(t) for t in lines[0].split())
    body = np.array([ln.split() for ln in lines[1:1 + nnz]])
    if nnz == 0:
        return (nr, nc), np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0)
    return (nr, nc), body[:, 0].astype(np.int64) - 1, body[:, 1].astype(np.int64) - 1, body[:, 2].astype(np.float64)


def write_mtx_csc(path, n_rows, n_cols, outer, inner, values):
    """Eigen::saveMarket layout of a column-major sparse matrix: header, sizes, then (row, col, value) column by column"""
    with open(path, "w") as fh:
        fh.write("%%MatrixMarket matrix coordinate  real general\n")
        fh.write(f"{n_rows} {n_cols} {len(values)}\n")
        for j in range(n_cols):
            for k in range(outer[j], outer[j + 1]):
                fh.write(f"{inner[k] + 1} {j + 1} {values[k]:.17g}\n")
