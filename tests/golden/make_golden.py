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
MESHES = ["c_shaped", "unit_square", "unit_square_16", "unit_square_32", "unit_sphere", "surface"]


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
        print(m, out[m + "/points"].shape, out[m + "/elements"].shape, int(out[m + "/boundary"].sum()),
              int(out[m + "/n_edges_file"]))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "meshes.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    sys.exit(main())
