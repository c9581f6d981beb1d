#!/usr/bin/env python
"""Benchmark of the FE assembly + sparse solve hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 119]

Workload (config.workload): BASELINE.json configs[3] -- 3D Laplacian, P1, synthetic unit-cube Kuhn mesh n=119
(1,728,000 nodes, 10,110,954 tetrahedra), stiffness assembly + CG to 1e-8.  One "step" = one full assembly of the
stiffness matrix (local-assembly kernel + segmented-reduction kernel) with mesh, dof table, pattern and scatter map
already resident in HBM.  `value` = tetrahedra assembled per second over all ranks.  The CG solve, the load vector
and the one-off setup (upload + pattern/scatter-map build) are timed separately and reported in the same line.
`e2e` = the same metric through the reference-facing call (Assembler(...).discretize_operator) with HOST buffers:
mesh upload, pattern build, assembly and CSC download all inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "elements assembled/s & CG solve s (3D P1 Laplacian 10M tets), 1-8 B200"
B_ASM_P1_TET = 16 + 96 + 40 + 20  # dof row + vertex coords + scatter map + CSC values
# dram__bytes_read.sum + dram__bytes_write.sum of one k_fused_assemble launch at n=119 (ncu --set full, round 1)
TRAFFIC_FUSED_BYTES = 960.1e6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm for the path.  The reference cannot be compiled here
    (Eigen 3.4 absent: DESIGN.md), so this times the oracle port on a bounded sample of the same workload.  The reference
    has no threading of any kind; to give the CPU arm every host thread it can use, the per-cell integration runs under
    OpenMP on all cores (bit-identical matrix, oracle/fdapde_oracle.c: orc_assemble_operator_mt) while setFromTriplets
    and the mirror pass stay serial as in Eigen.  The faithful 1-core number is the `cpu_baseline` of the default arm."""
    if rank != 0:
        return
    import __graft_entry__ as g
    fdb = g.load_package()
    from oracle import oracle as orc
    import scipy.sparse as sp
    ns = args.ref_n
    nodes, cells, bnd = fdb.meshes.unit_cube(ns)
    n = nodes.shape[0]
    cores = os.cpu_count() or 1
    times = []
    for k in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        o, i, v = orc.assemble_operator_mt(1, nodes, cells, cells, n, [(orc.LAPLACIAN, -1.0)], True, n_threads=cores)
        t = time.perf_counter() - t0
        if k >= args.warmup:
            times.append(t)
    t_step = float(np.mean(times))
    val = cells.shape[0] / t_step
    q = orc.quadrature_nodes(1, nodes, cells)
    b = orc.assemble_forcing(1, nodes, cells, cells, n, 3 * np.pi ** 2 * np.prod(np.sin(np.pi * q), axis=1))
    orc.set_dirichlet(o, i, v, bnd, np.zeros(n), b)
    Ar = sp.csc_matrix((v, i, o), shape=(n, n)).tocsr()
    Ar.sort_indices()
    t0 = time.perf_counter()
    _, iters, rel = orc.cg(Ar.indptr, Ar.indices, Ar.data, b, np.zeros(n), rtol=1e-8)
    t_cg = time.perf_counter() - t0
    sample = f"unit cube n={ns}: {cells.shape[0]} tets, {n} dofs (same mesh family as the n=119 workload)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "elements/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3D Laplacian P1, unit-cube Kuhn mesh n=119 (10,110,954 tets), stiffness assembly "
                                   "+ CG 1e-8; reference arm runs a bounded sample", "sample": sample},
            "cpu_baseline": {"value": val, "unit": "elements/s", "cores": cores, "kind": "port", "sample": sample,
                             "host_cores": os.cpu_count(),
                             "what": "oracle port, per-cell integration on all host cores (OpenMP), triplet merge serial"},
            "solve": {"seconds": t_cg, "iters": iters, "rel_resid": rel, "n_dofs": n, "what": "CPU CG (oracle), sample"},
            "e2e": {"value": val, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=119, help="cubes per edge (119 -> 10,110,954 tets)")
    ap.add_argument("--ref-n", type=int, default=40, help="cube size of the bounded CPU sample (384,000 tets)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-warmup-s", type=float, default=0.6, help="0 under ncu")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    fdb = g.load_package()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- synthetic workload (host) --------------------------------------------------------------------------------
    nodes_g, cells_g, bnd_g = fdb.meshes.unit_cube(args.n)
    n_total_cells, n_total_dofs = cells_g.shape[0], nodes_g.shape[0]
    if world > 1:
        # element partition: this rank owns a contiguous block of dof rows and assembles every cell touching them
        # (halo cells are recomputed by the neighbour: no communication in assembly)
        loc = fdb.partition.partition_p1(nodes_g, cells_g, bnd_g, rank, world)
        nodes, cells, bnd = loc.nodes, loc.cells, loc.boundary
        n_owned = loc.n_owned
    else:
        loc, nodes, cells, bnd, n_owned = None, nodes_g, cells_g, bnd_g, n_total_dofs
    del nodes_g, cells_g
    n_dofs = nodes.shape[0]            # local dofs (owned + halo)
    local_cells = cells.shape[0]
    mesh = fdb.Triangulation(nodes, cells, bnd)
    stream = torch.cuda.current_stream()
    op = -fdb.laplacian()
    comm = None
    if world > 1:
        def bcast(obj):
            box = [obj]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        comm = fdb.Comm(rank, world, bcast)

    # ---- one-off setup: upload + pattern / scatter map / fused plan (timed, reported separately) -------------------
    barrier()
    t0 = time.perf_counter()
    space = fdb.Space(mesh, 1, cells, n_dofs, bnd)
    space.set_stream(stream.cuda_stream)
    if loc is not None:
        space.set_dof0_rule(loc.own0 == 0)
    space.prepare(symmetric=True)   # pattern + scatter map + fused plan
    A = fdb.Matrix(space)
    A.assemble(op)
    torch.cuda.synchronize()
    setup_s = max_over_ranks(time.perf_counter() - t0)
    nnz = A.nnz()
    space.set_profiling(True)

    # ---- timed region: K assemblies, device resident ---------------------------------------------------------------
    # warm-up: at least W steps, and long enough (0.6 s) for the clocks to settle and for nvidia-smi (100 ms period)
    # to sample the GPU under this very load right up to and through the timed region
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_w, k_w = time.perf_counter(), 0
    while k_w < args.warmup or time.perf_counter() - t_w < args.min_warmup_s:
        A.assemble(op)
        k_w += 1
        if k_w % 8 == 0:
            torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        A.assemble(op)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    # per-kernel split of the last step (events recorded by the library on the same stream)
    t_k1, t_k2 = space.last_timings()
    fused = t_k2 < 0.02 * max(t_k1, 1e-9)
    value = n_total_cells / (ms_step * 1e-3)

    hbm, peak_src = peaks()
    asm_bytes = B_ASM_P1_TET * n_total_cells / world  # owned share only: recomputed halo cells earn nothing
    achieved = asm_bytes / (ms_step * 1e-3) / 1e9
    kname = ("k_fused_assemble<3,1,sym,lap> (local matrices in shared memory + in-order segment sums, one launch)"
             if fused else "k_local_assemble<3,1,sym,lap> + k_segmented_reduce<sym>")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": TRAFFIC_FUSED_BYTES if (fused and world == 1 and args.n == 119) else None, "kernel": kname,
                "algorithmic_bytes_per_launch": asm_bytes, "bytes_per_element": B_ASM_P1_TET,
                "peak_source": peak_src, "ms_kernel_1": t_k1, "ms_kernel_2": t_k2,
                "traffic_source": "ncu --set full, profiles/r01_ncu_summary.md" if fused else None}

    # ---- load vector --------------------------------------------------------------------------------------------------
    nq = space.n_quad
    q = space.quadrature_nodes()
    fq_host = 3 * np.pi ** 2 * np.prod(np.sin(np.pi * q), axis=1)  # f at the quadrature nodes (host)
    del q
    fq = fdb.Vector(local_cells * nq, fq_host)
    b = fdb.Vector(n_dofs)
    lib = fdb.lib()
    for _ in range(3):
        lib.fdb_assemble_forcing(space.h, fq.h, b.h)
    barrier()
    e0.record(stream)
    for _ in range(10):
        lib.fdb_assemble_forcing(space.h, fq.h, b.h)
    e1.record(stream)
    barrier()
    ms_force = max_over_ranks(e0.elapsed_time(e1) / 10)

    line = {"metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D Laplacian P1, unit-cube Kuhn mesh n={args.n} ({n_total_cells} tets, "
                                   f"{n_total_dofs} dofs), stiffness assembly + CG 1e-8 (BASELINE configs[3])",
                       "l2": "inputs larger than L2 (block cell lists + gather lists + values ~ 0.9 GB per assembly)",
                       "partition": "1 rank" if world == 1 else f"{world} row blocks, halo cells recomputed, "
                                                                 f"NCCL halo exchange + all-reduce in the solve"},
            "roofline": roofline, "clocks": clocks, "setup_s": setup_s, "ms_forcing": ms_force,
            "gpu_launches": (1 if fused else 2) * args.steps}

    # ---- solve: CG to 1e-8 (distributed when world > 1) ----------------------------------------------------------------
    if comm is not None:
        A.set_partition(comm, loc)
        if os.environ.get("FDB_PEER", "1") == "1":
            def gather(obj):
                out = [None] * world
                dist.all_gather_object(out, obj)
                return out
            A.enable_peer_memory(loc, gather)   # persistent CG: halo pushes + reductions over NVLink peer memory
    g_vec = fdb.Vector(n_dofs).fill(0.0)
    x = fdb.Vector(n_dofs).fill(0.0)
    A.set_dirichlet(g_vec, b, x)
    opts = fdb.SolverOptions("cg", rtol=1e-8, check_every=50)
    st = A.solve(b, x, opts)            # warm-up (allocates the workspace, NCCL channels)
    x.fill(0.0)
    barrier()
    st = A.solve(b, x, opts)
    t_solve = max_over_ranks(st["seconds"])
    it = max(st["iters"], 1)
    nnz_total = sum_over_ranks(nnz) if world > 1 else nnz   # includes the (incomplete) halo rows at N > 1
    b_cg = 12 * nnz_total + 92 * n_total_dofs
    line["solve"] = {"seconds": t_solve, "iters": st["iters"], "rel_resid": st["rel_resid"],
                     "converged": st["converged"], "us_per_iter": t_solve / it * 1e6,
                     "roofline": {"bound": "hbm", "achieved": b_cg * it / t_solve / 1e9, "peak": hbm * world,
                                  "unit": "GB/s", "frac": b_cg * it / t_solve / 1e9 / (hbm * world),
                                  "bytes_per_iter": b_cg,
                                  "kernel": ("CG iteration: k_spmv_sell<dot> + k_cg_update + k_cg_direction, CUDA-graph replay" if world == 1 else
                                            "k_cg_persistent: one cooperative kernel per rank, halo + reductions over NVLink peer memory")}}
    # SpMV alone (with its halo exchange at N > 1)
    y = fdb.Vector(n_dofs)
    for _ in range(5):
        A.spmv(x, y)
    barrier()
    e0.record(stream)
    for _ in range(50):
        A.spmv(x, y)
    e1.record(stream)
    barrier()
    ms_spmv = max_over_ranks(e0.elapsed_time(e1) / 50)
    b_spmv = 12 * nnz_total + 4 * (n_total_dofs + 1) + 16 * n_total_dofs
    line["spmv"] = {"ms": ms_spmv, "roofline": {"bound": "hbm", "achieved": b_spmv / (ms_spmv * 1e-3) / 1e9,
                                                "peak": hbm * world, "unit": "GB/s",
                                                "frac": b_spmv / (ms_spmv * 1e-3) / 1e9 / (hbm * world),
                                                "bytes_per_launch": b_spmv,
                                                "kernel": "k_spmv_sell (sliced-ELL, 16-bit column offsets)" if world == 1 else "k_spmv<4> + NCCL halo exchange"}}

    # ---- e2e: reference-facing call with host (pinned) buffers ------------------------------------------------------
    del A, space, x, y, b, fq, g_vec
    torch.cuda.synchronize()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    mesh_p = fdb.Triangulation(pin(np.asfortranarray(nodes).T).T, pin(cells), bnd)
    dofs_p = pin(np.asfortranarray(cells).T).T  # LagrangianBasis::dofs(): column-major, exists before the call
    times, h2d, d2h = [], 0, 0
    for k in range(1 + args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        asm = fdb.Assembler(mesh_p, 1, n_dofs, dofs_p)
        outer, inner, val = asm.discretize_operator(op)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        if k > 0:
            times.append(t)
        h2d = nodes.nbytes + cells.nbytes
        d2h = outer.nbytes + inner.nbytes + val.nbytes
        del asm
    e2e_s = max_over_ranks(float(np.mean(times))) if times else float("nan")
    line["e2e"] = {"value": n_total_cells / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": int(sum_over_ranks(h2d)),
                   "d2h_bytes_per_step": int(sum_over_ranks(d2h)), "seconds_per_step": e2e_s,
                   "what": "Assembler(mesh, ...).discretize_operator(-laplacian) from host arrays: upload + pattern "
                           "build + assembly + CSC download"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        ns = args.ref_n
        sn, sc, sb = fdb.meshes.unit_cube(ns)
        t0 = time.perf_counter()
        orc.assemble_operator(1, sn, sc, sc, sn.shape[0], [(orc.LAPLACIAN, -1.0)], True)
        t = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sc.shape[0] / t, "unit": "elements/s", "cores": 1, "kind": "port",
                                "host_cores": os.cpu_count(),
                                "sample": f"oracle (-O2 -march=x86-64, 1 thread like the reference) on unit cube "
                                          f"n={ns}: {sc.shape[0]} tets in {t:.1f} s"}
    if rank == 0:
        print(json.dumps(line))
    del comm
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
