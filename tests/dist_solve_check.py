"""Run under torchrun (one rank per GPU): distributed assembly + CG / BiCGSTAB through the C ABI, checked on rank 0
against the single-GPU solve of the same system and the oracle's direct solve.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tests/dist_solve_check.py [n]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    fdb = g.load_package()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))))
    n_cube = int(sys.argv[1]) if len(sys.argv) > 1 else 12

    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    comm = fdb.Comm(rank, world, bcast)
    nodes, cells, bnd = fdb.meshes.unit_cube(n_cube)
    nodes = fdb.meshes.jitter(nodes, bnd, 1.0 / n_cube)
    n = nodes.shape[0]
    fail = []
    for kind, expr in (("cg", -fdb.laplacian() + fdb.reaction(1.0)),
                       ("bicgstab", -fdb.laplacian() + fdb.advection([1.0, -0.5, 0.25]))):
        for jacobi in (False, True):
            loc = fdb.partition.partition_p1(nodes, cells, bnd, rank, world)
            mesh = fdb.Triangulation(loc.nodes, loc.cells, loc.boundary)
            nl = loc.local_to_global.size
            space = fdb.Space(mesh, 1, loc.cells, nl, loc.boundary)
            space.set_dof0_rule(loc.own0 == 0)  # only the owner of global dof 0 mirrors the reference's dof-0 quirk
            A = fdb.Matrix(space).assemble(expr)
            A.set_partition(comm, loc)
            if os.environ.get("FDB_PEER", "1") == "1" and kind == "cg":
                def gather(obj):
                    out = [None] * world
                    dist.all_gather_object(out, obj)
                    return out
                A.enable_peer_memory(loc, gather)   # persistent CG over NVLink peer memory (no NCCL in the loop)
            q = space.quadrature_nodes()
            f = np.prod(np.sin(np.pi * q), axis=1) + 1.0
            b = fdb.Vector(nl)
            fq = fdb.Vector(f.size, f)
            assert fdb.lib().fdb_assemble_forcing(space.h, fq.h, b.h) == 0
            gl = loc.nodes[:, 0] + 2 * loc.nodes[:, 1]
            x = fdb.Vector(nl).fill(0.0)
            A.set_dirichlet(fdb.Vector(nl, gl), b, x)
            st = A.solve(b, x, fdb.SolverOptions(kind, rtol=1e-11, jacobi=jacobi, check_every=10))
            xl = x.download()[:loc.n_owned]
            parts = [None] * world
            dist.all_gather_object(parts, (loc.own0, xl))
            if rank == 0:
                u = np.zeros(n)
                for own0, xp in parts:
                    u[own0:own0 + xp.size] = xp
                # single-GPU solve of the same problem
                m1 = fdb.Triangulation(nodes, cells, bnd)
                pde = fdb.PDE(m1, expr, 1, forcing=lambda qq: np.prod(np.sin(np.pi * qq), axis=1) + 1.0,
                              solver=fdb.SolverOptions(kind, rtol=1e-11, jacobi=jacobi))
                pde.set_dirichlet_bc(nodes[:, 0] + 2 * nodes[:, 1])
                pde.init()
                pde.solve()
                e1 = np.linalg.norm(u - pde.solution()) / np.linalg.norm(pde.solution())
                ok = st["converged"] and e1 < 1e-8
                print(f"[dist {world} ranks] {kind} jacobi={jacobi}: iters {st['iters']} (1 GPU: {pde.stats['iters']}) "
                      f"rel diff vs 1 GPU {e1:.2e} {'OK' if ok else 'FAIL'}", flush=True)
                if not ok:
                    fail.append((kind, jacobi, e1))
            del A, space
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK_PASS" if not fail else f"DIST_CHECK_FAIL {fail}", flush=True)
    sys.exit(1 if fail else 0)


if __name__ == "__main__":
    main()
