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
n = int(os.environ.get("AB_N", "119"))
reps = int(os.environ.get("AB_REPS", "30"))
nodes, cells, bnd = fdb.meshes.unit_cube(n)
mesh = fdb.Triangulation(nodes, cells, bnd)
stream = torch.cuda.current_stream()
space = fdb.Space(mesh, 1, cells, nodes.shape[0], bnd)
space.set_stream(stream.cuda_stream)
op = -fdb.laplacian()
A = fdb.Matrix(space)
A.assemble(op)            # two-kernel path (first assembly)
ref = A.download_csc()[2].copy()
space.prepare(symmetric=True)
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
print(f"{os.environ.get('FDB_LIB_PATH', 'default')}: bit-identical={same} median {np.median(ts):.4f} ms min {ts.min():.4f} ms "
      f"({cells.shape[0] / np.median(ts) / 1e6:.2f} G tets/s)")
