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
            if os.environ.get("FDB_PEER", "1") == "1":
                def gather(obj):
                    out = [None] * world
                    dist.all_gather_object(out, obj)
                    return out
                A.enable_peer_memory(loc, gather)   # persistent CG / BiCGSTAB over NVLink peer memory (no NCCL in the loop)
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
    # ---- P2 spaces: general dof tables (owner = rank of the lowest-id incident cell, SURVEY 8e "P2 caveat") -------------
    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    cases = [("2d", fdb.meshes.unit_square(20), "bicgstab", lambda: -fdb.laplacian() + fdb.advection([-1.0, 0.0]) + fdb.reaction(1.0)),
             ("2d", fdb.meshes.unit_square(20), "cg", lambda: -fdb.laplacian() + fdb.reaction(1.0)),
             # n = 8: with 2, 4 or 8 ranks the cell slabs are whole layers of cubes (two neighbours per rank at most);
             # FDB_DIST_N3D=5 gives slabs that cut through layers (3-5 neighbours per rank)
             ("3d", fdb.meshes.unit_cube(int(os.environ.get("FDB_DIST_N3D", "8"))), "cg",
              lambda: -fdb.laplacian() + fdb.reaction(1.0))]
    for tag, (nodes2, cells2, bnd2), kind, mk in cases:
        expr = mk()
        m1 = fdb.Triangulation(nodes2, cells2, bnd2)
        basis = fdb.LagrangianBasis(m1, 2)         # global dof table (every rank derives the same one)
        dofs, nd, bd = basis.dofs(), basis.size(), basis.boundary_dofs()
        loc = fdb.partition.partition_dofs(nodes2, cells2, dofs, nd, bd, rank, world)
        mesh = fdb.Triangulation(loc.nodes, loc.cells, np.asarray(bnd2).ravel()[loc.node_ids])
        nl = loc.n_local_dofs
        space = fdb.Space(mesh, 2, loc.dofs, nl, loc.boundary, pass_cells=True)   # vertex ids != dof ids locally
        space.set_dof0_rule(loc.owns_dof0)
        A = fdb.Matrix(space).assemble(expr)
        A.set_partition(comm, loc)
        if os.environ.get("FDB_PEER", "1") == "1":
            A.enable_peer_memory(loc, gather)
        # distributed SpMV (halo exchange) against the single-GPU matrix, before any boundary condition
        xr = np.random.default_rng(5).standard_normal(nd)
        yv = fdb.Vector(nl)
        A.spmv(fdb.Vector(nl, xr[loc.local_to_global]), yv)
        yparts = gather((loc.local_to_global[:loc.n_owned], yv.download()[:loc.n_owned]))
        if rank == 0:
            import scipy.sparse as sp
            sg = fdb.Space(m1, 2, dofs, nd, bd)
            o1, i1, v1 = fdb.Matrix(sg).assemble(expr).download_csc()
            yref = sp.csc_matrix((v1, i1, o1), shape=(nd, nd)) @ xr
            yd = np.full(nd, np.nan)
            for gid, yp in yparts:
                yd[gid] = yp
            es = np.max(np.abs(yd - yref)) / np.max(np.abs(yref))
            print(f"[dist {world} ranks] P2 {tag} {kind}: distributed SpMV rel err {es:.2e}", flush=True)
            if not es < 1e-12:
                fail.append((tag, kind, "spmv", es))
        q = space.quadrature_nodes()
        f = np.pi * np.sin(np.pi * q[:, 1])
        b = fdb.Vector(nl)
        fq = fdb.Vector(f.size, f)
        assert fdb.lib().fdb_assemble_forcing(space.h, fq.h, b.h) == 0
        xc = space.dofs_coords()
        gl = xc[:, 0] + 2 * xc[:, 1]
        x = fdb.Vector(nl).fill(0.0)
        A.set_dirichlet(fdb.Vector(nl, gl), b, x)
        rtol = 1e-11 if kind == "cg" else 1e-9
        st = A.solve(b, x, fdb.SolverOptions(kind, rtol=rtol, check_every=10), raise_on_fail=False)
        parts = gather((loc.local_to_global[:loc.n_owned], x.download()[:loc.n_owned]))
        if rank == 0:
            u = np.full(nd, np.nan)
            for gid, xp in parts:
                u[gid] = xp
            pde = fdb.PDE(m1, expr, 2, forcing=lambda qq: np.pi * np.sin(np.pi * qq[:, 1]),
                          solver=fdb.SolverOptions(kind, rtol=rtol))
            xg = pde.dof_coords()
            pde.set_dirichlet_bc(xg[:, 0] + 2 * xg[:, 1])
            pde.init()
            pde.solve()
            e1 = np.linalg.norm(u - pde.solution()) / np.linalg.norm(pde.solution())
            ok = st["converged"] and e1 < (1e-8 if kind == "cg" else 1e-6)
            print(f"[dist {world} ranks] P2 {tag} {kind}: iters {st['iters']} resid {st['rel_resid']:.2e} (1 GPU: "
                  f"{pde.stats['iters']} its, resid {pde.stats['rel_resid']:.2e}, ok {pde.success}) "
                  f"rel diff vs 1 GPU {e1:.2e} {'OK' if ok else 'FAIL'}", flush=True)
            if not ok:
                fail.append((tag, kind, e1))
        del A, space
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_CHECK_PASS" if not fail else f"DIST_CHECK_FAIL {fail}", flush=True)
    sys.exit(1 if fail else 0)


if __name__ == "__main__":
    main()
