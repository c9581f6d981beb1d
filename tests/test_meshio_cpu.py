"""Mesh / matrix file I/O in the reference's formats (utils/IO/csv_reader.h:75-118,
test/src/utils/mesh_loader.h:62-84).  CPU only; fixtures are written here in the layout of the reference's own files
(quoted header, quoted 1-based row names, 1-based element ids) from the committed golden meshes."""
import numpy as np
import pytest


def test_csv_reader_semantics(fdb, tmp_path):
    p = tmp_path / "a.csv"
    p.write_text('"","V1","V2","V3"\n"1","-0.925"," 1.5e-16",NA\n"2", 3 ,"4","nan"\n')
    a = fdb.meshio.read_csv(str(p))
    assert a.shape == (2, 3)
    assert a[0, 0] == -0.925 and a[0, 1] == 1.5e-16 and np.isnan(a[0, 2])
    assert a[1, 0] == 3.0 and a[1, 1] == 4.0 and np.isnan(a[1, 2])
    q = tmp_path / "b.csv"
    q.write_text('"","V1"\n"1","7"\n"2","0"\n')
    b = fdb.meshio.read_csv(str(q), np.int32)
    assert b.dtype == np.int32 and b.ravel().tolist() == [7, 0]


@pytest.mark.parametrize("name", ["c_shaped", "unit_sphere", "surface"])
def test_mesh_directory_round_trip(fdb, golden_meshes, tmp_path, name):
    pts, els, bnd = golden_meshes(name)
    d = str(tmp_path / name)
    fdb.meshio.save_mesh(d, pts, els, bnd)
    first = open(d + "/elements.csv").read().splitlines()[:2]
    assert first[0].startswith('"","V1"') and first[1].split(",")[1] == str(els[0, 0] + 1)   # 1-based on disk
    m = fdb.meshio.load_mesh(d)
    assert np.array_equal(m["points"], pts)          # %.17g round-trips fp64 exactly
    assert np.array_equal(m["elements"], els) and m["elements"].dtype == np.int32
    assert np.array_equal(m["boundary"], bnd)


def test_matrix_market_round_trip(fdb, golden_meshes, tmp_path):
    from oracle import oracle as orc
    pts, els, bnd = golden_meshes("unit_square_16")
    o, i, v = orc.assemble_operator(1, pts, els, els, pts.shape[0], [(orc.LAPLACIAN, -1.0)], True)
    p = str(tmp_path / "stiff.mtx")
    fdb.meshio.write_mtx_csc(p, pts.shape[0], pts.shape[0], o, i, v)
    shape, r, c, vals = fdb.meshio.read_mtx(p)
    assert shape == (pts.shape[0], pts.shape[0])
    assert np.array_equal(r, i) and np.array_equal(c, np.repeat(np.arange(pts.shape[0]), np.diff(o)))
    assert np.array_equal(vals, v)
