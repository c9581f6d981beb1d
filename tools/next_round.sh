# Development helper: first GPU call of the next round -- measures the switches that were finished after the round-1
# GPU budget was spent.  One GPU, about a minute:
#   gpurun --timeout 600 -- 'bash tools/next_round.sh'
set -x
# 1. bank-aware cell numbering of the fused plan: correctness gate, then A/B timing
FDB_FUSED_BANKS=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_and or ladder or operators" 2>&1 | tail -3
timeout 200 python tools/ab_assembly.py 2>&1 | tail -1
FDB_FUSED_BANKS=1 timeout 200 python tools/ab_assembly.py 2>&1 | tail -1
# 2. surface point location / evaluation kernels
FDB_TEST_SURFACE_EVAL=1 timeout 100 python -m pytest tests/test_gpu_evaluation.py -x -q -m gpu -k surface 2>&1 | tail -3
# 3. single-reduction CG on one GPU (persistent kernel): iterations and time against the default loop
timeout 200 python - <<'PY'
import os, time, numpy as np
import __graft_entry__ as g
fdb = g.load_package()
nodes, cells, bnd = fdb.meshes.unit_cube(119)
n = nodes.shape[0]
s = fdb.Space(fdb.Triangulation(nodes, cells, bnd), 1, cells, n, bnd)
A = fdb.Matrix(s).assemble(-fdb.laplacian())
q = s.quadrature_nodes()
f = 3 * np.pi ** 2 * np.prod(np.sin(np.pi * q), axis=1)
b = fdb.Vector(n); fq = fdb.Vector(f.size, f)
assert fdb.lib().fdb_assemble_forcing(s.h, fq.h, b.h) == 0
x = fdb.Vector(n).fill(0.0)
A.set_dirichlet(fdb.Vector(n).fill(0.0), b, x)
for mode, env in (("multi-kernel", {}), ("persistent", {"FDB_PERSISTENT": "1"}), ("persistent cg1", {"FDB_PERSISTENT": "1", "FDB_CG1": "1"})):
    for k in ("FDB_PERSISTENT", "FDB_CG1"):
        os.environ.pop(k, None)
    os.environ.update(env)
    fdb.lib().fdb_set_persistent_cg(2 if "FDB_PERSISTENT" in env else 0)
    for rep in range(2):
        x.fill(0.0)
        st = A.solve(b, x, fdb.SolverOptions("cg", rtol=1e-8))
    print(mode, st)
PY
