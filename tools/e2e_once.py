"""Development helper: the reference-facing call of the bench's e2e block, three times (for a launch list under ncu)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import torch
fdb = g.load_package()
nodes, cells, bnd = fdb.meshes.unit_cube(int(os.environ.get("AB_N", "119")))
n = nodes.shape[0]
nodes_p = fdb.api.pinned_copy(np.asfortranarray(nodes).T).T
cells_p = fdb.api.pinned_copy(cells)
dofs_p = fdb.api.pinned_copy(np.asfortranarray(cells).T).T
mesh = fdb.Triangulation(nodes_p, cells_p, bnd)
op = -fdb.laplacian()
for k in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    asm = fdb.Assembler(mesh, 1, n, dofs_p)
    t1 = time.perf_counter()
    o, i, v = asm.discretize_operator(op)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"call {k}: create {1e3*(t1-t0):.2f} ms, discretize {1e3*(t2-t1):.2f} ms", flush=True)
    del asm, o, i, v
