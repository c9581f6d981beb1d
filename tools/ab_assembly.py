"""Development helper: times the fused stiffness assembly of workload C4 for one build of the library
(FDB_LIB_PATH selects a variant built with `build.py --variant=...`) and checks it bit-for-bit against the
two-kernel path.  Prints one line.  Not part of the product or of the bench contract."""
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g

fdb = g.load_package()
cfg = os.environ.get("AB_CONFIG", "c4")     # c4: 3D P1 stiffness, c2: 2D P1 stiffness, c3: 2D P2 ADR (non-symmetric)
reps = int(os.environ.get("AB_REPS", "30"))
stream = torch.cuda.current_stream()
if cfg == "c4":
    n = int(os.environ.get("AB_N", "119"))
    nodes, cells, bnd = fdb.meshes.unit_cube(n)
    mesh = fdb.Triangulation(nodes, cells, bnd)
    space = fdb.Space(mesh, 1, cells, nodes.shape[0], bnd)
    op = {"mass": fdb.reaction(1.0), "adr": -fdb.laplacian() + fdb.advection([1.0, -0.5, 0.25]) + fdb.reaction(2.0),
          "diff": -fdb.diffusion([[2.0, 0.3, 0.0], [0.3, 1.0, 0.1], [0.0, 0.1, 1.5]])}.get(os.environ.get("AB_OP"), -fdb.laplacian())
elif cfg == "c2":
    nodes, cells, bnd = fdb.meshes.unit_square(int(os.environ.get("AB_N", "1414")))
    mesh = fdb.Triangulation(nodes, cells, bnd)
    space = fdb.Space(mesh, 1, cells, nodes.shape[0], bnd)
    op = fdb.reaction(1.0) if os.environ.get("AB_OP") == "mass" else -fdb.laplacian()
elif cfg == "p2tet":   # C5-like: 3D P2 on a cube that matches one slab of the 8-way C5 partition in size (n=76: 2.63 M tets)
    nodes, cells, bnd = fdb.meshes.unit_cube(int(os.environ.get("AB_N", "76")))
    mesh = fdb.Triangulation(nodes, cells, bnd)
    basis = fdb.LagrangianBasis(mesh, 2)
    space = fdb.Space(mesh, 2, basis.dofs(), basis.size(), basis.boundary_dofs())
    op = fdb.reaction(1.0) if os.environ.get("AB_OP") == "mass" else -fdb.laplacian()
else:
    nodes, cells, bnd = fdb.meshes.unit_square(int(os.environ.get("AB_N", "1000")))
    mesh = fdb.Triangulation(nodes, cells, bnd)
    basis = fdb.LagrangianBasis(mesh, 2)
    space = fdb.Space(mesh, 2, basis.dofs(), basis.size(), basis.boundary_dofs())
    op = -fdb.laplacian() + fdb.advection([-1.0, 0.0]) + fdb.reaction(1.0)
space.set_stream(stream.cuda_stream)
A = fdb.Matrix(space)
A.assemble(op)            # two-kernel path (first assembly)
ref = A.download_csc()[2].copy()
space.set_profiling(True)
space.set_fused(False)
A.assemble(op)
torch.cuda.synchronize()
t2k = space.last_timings()
space.set_fused(True)
space.prepare(symmetric=op.is_symmetric)
for _ in range(5):
    A.assemble(op)
got = A.download_csc()[2]
same = bool(np.array_equal(ref, got))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    A.assemble(op)
    e1.record(stream)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = np.array(ts)
print(f"{cfg} {os.environ.get('AB_OP', '')} {os.environ.get('FDB_LIB_PATH', 'default')}: fused={space.last_path()[0]} two-kernel {t2k[0]:.3f}+{t2k[1]:.3f} ms; bit-identical={same} median {np.median(ts):.4f} ms min {ts.min():.4f} ms "
      f"({cells.shape[0] / np.median(ts) / 1e6:.2f} G cells/s)")
