#!/usr/bin/env python
"""Benchmark of the FE assembly + sparse solve hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n 119]

Workload (config.workload): BASELINE.json configs[3] (C4) -- 3D Laplacian, P1, synthetic unit-cube Kuhn mesh n=119
(1,728,000 nodes, 10,110,954 tetrahedra), stiffness assembly + CG to 1e-8.  One "step" = one full assembly of the
stiffness matrix with mesh, dof table, pattern and plan already resident in HBM.  `value` = tetrahedra assembled per
second over all ranks.  The CG solve, the SpMV, the load vector and the one-off setup are timed separately and reported
in the same line, and so are the other BASELINE configurations that fit the launch (`configs`: C2 at N=1, C3 at every
N, C5 at N=8), each with its own roofline from SURVEY.md 8(d)'s bytes.
`e2e` = the same metric through the reference-facing call (Assembler(...).discretize_operator) with HOST buffers: mesh
upload, pattern build, assembly and CSC download all inside the timed region, with a per-stage breakdown.
`parity` = checks made in this very run (solution against the exact one and, at N>1, against rank 0's single-GPU solve;
owned matrix rows of every rank bit-compared with the single-GPU rows).
"""
import argparse
import hashlib
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
# dof row + vertex coords + scatter map + CSC values, bytes per element and matrix
B_ASM = {"c2": 112, "c3": 400, "c4": 172, "c5": 663}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture of the same kernel on
    the same workload (profiles/traffic.json, written by profiles/make_traffic.py from the raw ncu exports).  None when
    no capture of this kernel / workload is committed: the number is never invented."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    for row in json.load(open(p)):
        if row["kernel"] == kernel and row["workload"] == workload:
            return row["dram_bytes"], row["source"]
    return None, None


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


def c4_workload_text(n, n_cells, n_dofs):
    return (f"3D Laplacian P1, unit-cube Kuhn mesh n={n} ({n_cells} tets, {n_dofs} dofs), stiffness assembly "
            f"+ CG 1e-8 (BASELINE configs[3])")


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm for the path, on the SAME workload (n=119, 10.1 M tets).
    The reference cannot be compiled here (Eigen 3.4 absent: DESIGN.md), so this times the oracle port.  The reference
    has no threading of any kind; to give the CPU arm every host thread it can use, the per-cell integration runs under
    OpenMP on all cores (bit-identical matrix, oracle/fdapde_oracle.c: orc_assemble_operator_mt) while setFromTriplets
    and the mirror pass stay serial as in Eigen.  One step = one full assembly.  If the projected run (K + W steps)
    would not end within --ref-budget-s, the mesh is shrunk (and the line says so); the faithful 1-core number is the
    `cpu_baseline` of the default arm."""
    if rank != 0:
        return
    import __graft_entry__ as g
    fdb = g.load_package()
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    ns = args.ref_n if args.ref_n > 0 else args.n
    nodes, cells, bnd = fdb.meshes.unit_cube(ns)
    n = nodes.shape[0]
    op = [(orc.LAPLACIAN, -1.0)]

    def one():
        t0 = time.perf_counter()
        r = orc.assemble_operator_mt(1, nodes, cells, cells, n, op, True, n_threads=cores)
        return time.perf_counter() - t0, r

    t_first, res = one()                              # doubles as the first warm-up step
    total_steps = args.warmup + args.steps
    if t_first * total_steps > args.ref_budget_s and args.ref_n <= 0:
        ns = max(8, int(ns * (args.ref_budget_s / (t_first * total_steps)) ** (1.0 / 3.0)))
        nodes, cells, bnd = fdb.meshes.unit_cube(ns)
        n = nodes.shape[0]
        t_first, res = one()
    times = []
    for k in range(1, total_steps):
        t, res = one()
        if k >= args.warmup:
            times.append(t)
    if not times:
        times = [t_first]
    t_step = float(np.mean(times))
    val = cells.shape[0] / t_step
    same = ns == args.n
    sample = (f"full workload: unit cube n={ns}, {cells.shape[0]} tets, {n} dofs, one whole assembly per step"
              if same else f"unit cube n={ns}: {cells.shape[0]} tets, {n} dofs (same mesh family as the n={args.n} "
                           f"workload, shrunk to fit {args.ref_budget_s:.0f} s)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "elements/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": c4_workload_text(args.n, 6 * args.n ** 3, (args.n + 1) ** 3),
                       "same_config": same, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "elements/s", "cores": cores, "kind": "port", "sample": sample,
                             "host_cores": os.cpu_count(),
                             "what": "oracle port, per-cell integration on all host cores (OpenMP), triplet merge serial"},
            "e2e": {"value": val, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_ref_solve:
        import scipy.sparse as sp
        o, i, v = res
        q = orc.quadrature_nodes(1, nodes, cells)
        b = orc.assemble_forcing(1, nodes, cells, cells, n, 3 * np.pi ** 2 * np.prod(np.sin(np.pi * q), axis=1))
        orc.set_dirichlet(o, i, v, bnd, np.zeros(n), b)
        Ar = sp.csc_matrix((v, i, o), shape=(n, n)).tocsr()
        Ar.sort_indices()
        t0 = time.perf_counter()
        _, iters, rel = orc.cg(Ar.indptr, Ar.indices, Ar.data, b, np.zeros(n), rtol=1e-8)
        line["solve"] = {"seconds": time.perf_counter() - t0, "iters": iters, "rel_resid": rel, "n_dofs": n,
                         "what": "CPU CG (oracle, 1 thread), same stopping rule"}
    print(json.dumps(line))


# =====================================================================================================================
class Ctx:
    """torch / torch.distributed plumbing shared by the workloads."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import __graft_entry__ as g
        self.torch, self.dist = torch, dist
        self.fdb = g.load_package()
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream()
        self.comm = None
        if self.world > 1:
            self.comm = self.fdb.Comm(self.rank, self.world, self.bcast)
        self.hbm, self.peak_src = peaks()

    def bcast(self, obj):
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def gather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _red(self, x, op):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX)

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM)

    def events(self):
        return self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)

    def time_loop(self, fn, reps, warm=3):
        """ms per call of fn(), CUDA events on the launching stream, max over ranks."""
        for _ in range(warm):
            fn()
        self.barrier()
        e0, e1 = self.events()
        e0.record(self.stream)
        for _ in range(reps):
            fn()
        e1.record(self.stream)
        self.barrier()
        return self.max(e0.elapsed_time(e1) / reps)

    def roof(self, nbytes, ms, kernel, **extra):
        ach = nbytes / (ms * 1e-3) / 1e9
        peak = self.hbm * self.world
        # achieved = ALGORITHMIC bytes / time: the accounting charges lists the kernels never read (scatter
        # map, 32-bit columns), so a kernel that moves fewer bytes than the accounting can exceed frac = 1 (SpMV, C2)
        d = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "kernel": kernel,
             "accounting": "algorithmic bytes per SURVEY 8(d), not DRAM traffic"}
        d.update(extra)
        return d


def gathered_solution(ctx, loc, x, n_global):
    """Global solution vector on every rank from the owned parts (reference numbering)."""
    parts = ctx.gather((loc.local_to_global[:loc.n_owned], x.download()[:loc.n_owned]))
    u = np.full(n_global, np.nan)
    for gid, xp in parts:
        u[gid] = xp
    return u


def owned_rows_digest(outer, inner, val, n_owned, l2g, n_global):
    """sha1 of the owned rows (global column ids ascending + values) of a local CSR/CSC matrix."""
    import scipy.sparse as sp
    end = int(outer[n_owned])
    m = sp.csr_matrix((val[:end], l2g[inner[:end]].astype(np.int64), outer[:n_owned + 1]), shape=(n_owned, n_global))
    m.sort_indices()
    h = hashlib.sha1()
    h.update(np.diff(m.indptr).astype(np.int32).tobytes())
    h.update(m.indices.astype(np.int32).tobytes())
    h.update(m.data.tobytes())
    return h.hexdigest()


def rows_digest_global(outer, inner, val, rows, n_global):
    import scipy.sparse as sp
    m = sp.csr_matrix((val, inner, outer), shape=(n_global, n_global))[rows]
    m.sort_indices()
    h = hashlib.sha1()
    h.update(np.diff(m.indptr).astype(np.int32).tobytes())
    h.update(m.indices.astype(np.int32).tobytes())
    h.update(m.data.tobytes())
    return h.hexdigest()


# ---- C4: the north-star workload (3D P1 Laplacian, 10.1 M tets) ---------------------------------------------------------
def bench_c4(ctx, args):
    fdb, torch = ctx.fdb, ctx.torch
    world, rank = ctx.world, ctx.rank
    nodes_g, cells_g, bnd_g = fdb.meshes.unit_cube(args.n)
    n_total_cells, n_total_dofs = cells_g.shape[0], nodes_g.shape[0]
    t_part = 0.0
    if world > 1:
        # element partition: this rank owns a contiguous block of dof rows and assembles every cell touching them
        # (halo cells are recomputed by the neighbour: no communication in assembly)
        t0 = time.perf_counter()
        loc = fdb.partition.partition_p1(nodes_g, cells_g, bnd_g, rank, world)
        t_part = time.perf_counter() - t0
        nodes, cells, bnd = loc.nodes, loc.cells, loc.boundary
    else:
        loc, nodes, cells, bnd = None, nodes_g, cells_g, bnd_g
    n_dofs = nodes.shape[0]            # local dofs (owned + halo)
    local_cells = cells.shape[0]
    mesh = fdb.Triangulation(nodes, cells, bnd)
    op = -fdb.laplacian()

    # ---- one-off setup: upload + pattern / scatter map / fused plan (timed, reported separately) -------------------
    ctx.barrier()
    t0 = time.perf_counter()
    space = fdb.Space(mesh, 1, cells, n_dofs, bnd)
    space.set_stream(ctx.stream.cuda_stream)
    if loc is not None:
        space.set_dof0_rule(loc.own0 == 0)
    space.prepare(symmetric=True)   # pattern + scatter map + fused plan
    A = fdb.Matrix(space)
    A.assemble(op)
    torch.cuda.synchronize()
    setup_s = ctx.max(time.perf_counter() - t0)
    nnz = A.nnz()

    # ---- timed region: K assemblies, device resident ---------------------------------------------------------------
    # warm-up: at least W steps, and long enough (0.6 s) for the clocks to settle and for nvidia-smi (100 ms period)
    # to sample the GPU under this very load right up to and through the timed region
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    t_w, k_w = time.perf_counter(), 0
    while k_w < args.warmup or time.perf_counter() - t_w < args.min_warmup_s:
        A.assemble(op)
        k_w += 1
        if k_w % 8 == 0:
            torch.cuda.synchronize()
    ctx.barrier()
    e0, e1 = ctx.events()
    e0.record(ctx.stream)
    for _ in range(args.steps):
        A.assemble(op)
    e1.record(ctx.stream)
    ctx.barrier()
    clocks = sampler.stop()
    ms_step = ctx.max(e0.elapsed_time(e1) / args.steps)
    # per-kernel split: one more step with the library's own events around each kernel (kept out of the timed region:
    # the three event records cost a few microseconds per step, which matters at N = 8 where a step is ~70 us)
    space.set_profiling(True)
    A.assemble(op)
    torch.cuda.synchronize()
    t_k1, t_k2 = space.last_timings()
    space.set_profiling(False)
    fused, launches = space.last_path()        # which path ran is reported by the library, not guessed from timings
    value = n_total_cells / (ms_step * 1e-3)

    asm_bytes = B_ASM["c4"] * n_total_cells   # whole job; recomputed halo cells earn nothing
    which = space.last_kernel()               # reported by the library: 2 persistent fused, 1 fused, 0 two kernels
    kshort = {2: "k_fused_persist<3,1,0>", 1: "k_fused_assemble<3,1,1,0>",
              0: "k_local_assemble<3,1,1,0>+k_segmented_reduce<1>"}[which]
    kname = {2: "k_fused_persist<M=3,sym,lean> (persistent CTAs; node coordinates and block lists prefetched by the bulk-copy "
                "engine one block ahead; local matrices in shared memory + in-order segment sums, one launch)",
             1: "k_fused_assemble<M=3,R=1,sym,lean> (local matrices in shared memory + in-order segment sums, one launch)",
             0: "k_local_assemble<3,1,sym,lean> + k_segmented_reduce<sym>"}[which]
    traffic, traffic_src = (measured_traffic(kshort, f"c4 n={args.n}") if world == 1 else (None, None))
    roofline = ctx.roof(asm_bytes, ms_step, kname, traffic=traffic, traffic_source=traffic_src,
                        algorithmic_bytes_per_launch=asm_bytes / world, bytes_per_element=B_ASM["c4"],
                        peak_source=ctx.peak_src, ms_kernel_1=t_k1, ms_kernel_2=t_k2,
                        # one launch alone between two events (no overlap of its tail with the next launch's head, which
                        # back-to-back launches of the persistent kernel enjoy): the conservative reading of the same kernel
                        frac_isolated_launch=(asm_bytes / world) / (max(t_k1 + t_k2, 1e-9) * 1e-3) / 1e9 / ctx.hbm,
                        path={2: "fused-persistent", 1: "fused", 0: "two-kernel"}[which])

    # ---- load vector --------------------------------------------------------------------------------------------------
    nq = space.n_quad
    q = space.quadrature_nodes()
    fq_host = 3 * np.pi ** 2 * np.prod(np.sin(np.pi * q), axis=1)  # f at the quadrature nodes (host)
    del q
    fq = fdb.Vector(local_cells * nq, fq_host)
    b = fdb.Vector(n_dofs)
    lib = fdb.lib()
    ms_force = ctx.time_loop(lambda: lib.fdb_assemble_forcing(space.h, fq.h, b.h), 10)

    line = {"metric": METRIC, "value": value, "unit": "elements/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": c4_workload_text(args.n, n_total_cells, n_total_dofs),
                       "l2": "inputs larger than L2 (block cell lists + gather lists + values ~ 0.9 GB per assembly)",
                       "partition": "1 rank" if world == 1 else f"{world} row blocks, halo cells recomputed, "
                                                                 f"halo exchange + reductions over NVLink peer memory"},
            "roofline": roofline, "clocks": clocks, "setup_s": setup_s, "ms_forcing": ms_force,
            "gpu_launches": launches * args.steps}

    # ---- matrix parity across ranks (before the Dirichlet rows change the values) ------------------------------------
    parity = {}
    if world > 1:
        o_l, i_l, v_l = A.download_csc()       # symmetric operator: CSC arrays == CSR arrays
        digest = owned_rows_digest(o_l, i_l, v_l, loc.n_owned, loc.local_to_global, n_total_dofs)
        digests = ctx.gather((loc.own0, loc.n_owned, digest))
        del o_l, i_l, v_l

    # ---- solve: CG to 1e-8 (distributed when world > 1) ----------------------------------------------------------------
    if ctx.comm is not None:
        A.set_partition(ctx.comm, loc)
        if os.environ.get("FDB_PEER", "1") == "1":
            A.enable_peer_memory(loc, ctx.gather)   # persistent CG: halo pushes + reductions over NVLink peer memory
    g_vec = fdb.Vector(n_dofs).fill(0.0)
    x = fdb.Vector(n_dofs).fill(0.0)
    A.set_dirichlet(g_vec, b, x)
    opts = fdb.SolverOptions("cg", rtol=1e-8, check_every=50)
    st = A.solve(b, x, opts)            # warm-up (allocates the workspace, peer channels)
    x.fill(0.0)
    ctx.barrier()
    st = A.solve(b, x, opts)
    t_solve = ctx.max(st["seconds"])
    it = max(st["iters"], 1)
    nnz_total = ctx.sum(nnz) if world > 1 else nnz   # includes the (incomplete) halo rows at N > 1
    b_cg = 12 * nnz_total + 92 * n_total_dofs
    line["solve"] = {"seconds": t_solve, "iters": st["iters"], "rel_resid": st["rel_resid"],
                     "converged": st["converged"], "us_per_iter": t_solve / it * 1e6,
                     "roofline": ctx.roof(b_cg * it, t_solve * 1e3,
                                          "CG iteration: k_spmv_sell<dot> + k_cg_update + k_cg_direction, CUDA-graph replay"
                                          if world == 1 else
                                          "k_cg1_sell: single-reduction CG, one cooperative sliced-ELL kernel per rank, halo + "
                                          "reductions over NVLink peer memory",
                                          bytes_per_iter=b_cg)}
    # solution parity: exact solution of the manufactured problem; at N > 1 also rank 0's own single-GPU solve
    u_loc = x.download()
    if world == 1:
        u = u_loc
    else:
        u = gathered_solution(ctx, loc, x, n_total_dofs)
    u_ex = np.prod(np.sin(np.pi * nodes_g), axis=1)
    parity["max_err_vs_exact"] = float(np.max(np.abs(u - u_ex)))   # O(h^2) discretisation error, h = 1/n
    parity["max_err_bound"] = 2.0 * (np.pi / args.n) ** 2
    parity["solution_ok"] = bool(parity["max_err_vs_exact"] < parity["max_err_bound"])

    # SpMV alone (with its halo exchange at N > 1)
    y = fdb.Vector(n_dofs)
    ms_spmv = ctx.time_loop(lambda: A.spmv(x, y), 50, warm=5)
    b_spmv = 12 * nnz_total + 4 * (n_total_dofs + 1) + 16 * n_total_dofs
    line["spmv"] = {"ms": ms_spmv,
                    "roofline": ctx.roof(b_spmv, ms_spmv, "k_spmv_sell (sliced-ELL, 16-bit column offsets)" +
                                         ("" if world == 1 else " + halo exchange"), bytes_per_launch=b_spmv)}

    if world > 1:
        # rank 0 repeats the whole problem on its one GPU: solution and owned rows of every rank against it
        ok_rows, rel = None, None
        if rank == 0:
            s1 = fdb.Space(fdb.Triangulation(nodes_g, cells_g, bnd_g), 1, cells_g, n_total_dofs, bnd_g)
            A1 = fdb.Matrix(s1).assemble(op)
            o1, i1, v1 = A1.download_csc()
            ok_rows = all(rows_digest_global(o1, i1, v1, slice(own0, own0 + no), n_total_dofs) == dg
                          for own0, no, dg in digests)
            del o1, i1, v1
            q1 = s1.quadrature_nodes()
            f1 = fdb.Vector(q1.shape[0], 3 * np.pi ** 2 * np.prod(np.sin(np.pi * q1), axis=1))
            del q1
            b1, x1 = fdb.Vector(n_total_dofs), fdb.Vector(n_total_dofs).fill(0.0)
            lib.fdb_assemble_forcing(s1.h, f1.h, b1.h)
            A1.set_dirichlet(fdb.Vector(n_total_dofs).fill(0.0), b1, x1)
            st1 = A1.solve(b1, x1, opts)
            u1 = x1.download()
            rel = float(np.linalg.norm(u - u1) / np.linalg.norm(u1))
            parity["single_gpu_iters"] = st1["iters"]
            del A1, s1, f1, b1, x1
        ok_rows, rel = ctx.bcast((ok_rows, rel))
        parity["owned_rows_bitwise_equal"] = ok_rows
        parity["rel_diff_vs_single_gpu_solution"] = rel
        parity["solution_ok"] = bool(parity["solution_ok"] and rel is not None and rel < 1e-8)
    line["parity"] = parity

    del A, space, x, y, b, fq, g_vec
    torch.cuda.synchronize()
    line["e2e"] = bench_e2e(ctx, args, nodes, cells, bnd, n_dofs, n_total_cells, op, t_part)
    return line


def bench_e2e(ctx, args, nodes, cells, bnd, n_dofs, n_total_cells, op, t_part):
    """The reference-facing call with host buffers: Assembler(mesh, integrator, n_dofs, dofs).discretize_operator(op)
    (fem_assembler.h:46-52).  Inputs are pinned host arrays, outputs land in pinned host arrays; upload, pattern build,
    assembly and download are all inside the timed region.  The stage breakdown comes from one extra step that
    synchronises between the stages."""
    fdb, torch = ctx.fdb, ctx.torch
    nodes_p = fdb.api.pinned_copy(np.asfortranarray(nodes).T).T        # column-major n_nodes x N, page-locked
    cells_p = fdb.api.pinned_copy(cells)
    dofs_p = fdb.api.pinned_copy(np.asfortranarray(cells).T).T         # LagrangianBasis::dofs(): column-major
    mesh_p = fdb.Triangulation(nodes_p, cells_p, bnd)
    times, h2d, d2h = [], 0, 0
    for k in range(1 + args.e2e_steps):
        ctx.barrier()
        t0 = time.perf_counter()
        asm = fdb.Assembler(mesh_p, 1, n_dofs, dofs_p)
        outer, inner, val = asm.discretize_operator(op)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        if k > 0:
            times.append(t)
        h2d = nodes.nbytes + cells.nbytes
        d2h = outer.nbytes + inner.nbytes + val.nbytes
        del asm, outer, inner, val
    e2e_s = ctx.max(float(np.mean(times))) if times else float("nan")
    # stage breakdown (one more step, synchronised between the stages)
    ctx.barrier()
    t0 = time.perf_counter()
    asm = fdb.Assembler(mesh_p, 1, n_dofs, dofs_p)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    asm.space.prepare_pattern(True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    A = fdb.Matrix(asm.space).assemble(op)
    asm.space.sync()
    t3 = time.perf_counter()
    o, i, v = A.download_csc(pinned=True)
    t4 = time.perf_counter()
    del A, asm, o, i, v
    br = {"upload_ms": (t1 - t0) * 1e3, "pattern_ms": (t2 - t1) * 1e3, "assemble_ms": (t3 - t2) * 1e3,
          "download_ms": (t4 - t3) * 1e3}
    out = {"value": n_total_cells / e2e_s, "unit": "elements/s", "h2d_bytes_per_step": int(ctx.sum(h2d)),
           "d2h_bytes_per_step": int(ctx.sum(d2h)), "seconds_per_step": e2e_s,
           "breakdown": {k: ctx.max(v) for k, v in br.items()},
           "what": "Assembler(mesh, ...).discretize_operator(-laplacian) from pinned host arrays into pinned host "
                   "arrays: upload + pattern build + assembly + CSC download"}
    if ctx.world > 1:
        out["not_timed"] = {"host_partition_s": ctx.max(t_part),
                            "what": "numpy row-block partition of the global mesh (mesh distribution, done once)"}
    return out


# ---- C2: 2D Poisson P1, 4 M triangles, stiffness + mass + CG (1 GPU) ---------------------------------------------------
def bench_c2(ctx, args):
    fdb = ctx.fdb
    N = 1414
    nodes, cells, bnd = fdb.meshes.unit_square(N)
    n = nodes.shape[0]
    s = fdb.Space(fdb.Triangulation(nodes, cells, bnd), 1, cells, n, bnd)
    s.set_stream(ctx.stream.cuda_stream)
    s.prepare(True)
    K, M = fdb.Matrix(s), fdb.Matrix(s)
    stiff, mass = -fdb.laplacian(), fdb.reaction(1.0)
    ms_k = ctx.time_loop(lambda: K.assemble(stiff), 20)
    fused, _ = s.last_path()
    persist = s.last_kernel() == 2
    ms_m = ctx.time_loop(lambda: M.assemble(mass), 20)
    q = s.quadrature_nodes()
    f = 2 * np.pi ** 2 * np.sin(np.pi * q[:, 0]) * np.sin(np.pi * q[:, 1])
    del q
    b, fq = fdb.Vector(n), fdb.Vector(f.size, f)
    fdb.lib().fdb_assemble_forcing(s.h, fq.h, b.h)
    x = fdb.Vector(n).fill(0.0)
    K.set_dirichlet(fdb.Vector(n).fill(0.0), b, x)
    opts = fdb.SolverOptions("cg", rtol=1e-8, check_every=50)
    K.solve(b, x, opts)
    x.fill(0.0)
    st = K.solve(b, x, opts)
    err = x.download() - np.sin(np.pi * nodes[:, 0]) * np.sin(np.pi * nodes[:, 1])
    nnz = K.nnz()
    it = max(st["iters"], 1)
    nb = B_ASM["c2"] * cells.shape[0]
    return {"workload": f"2D Poisson P1, unit square N={N} ({cells.shape[0]} triangles, {n} dofs), stiffness + mass "
                        f"+ CG 1e-8 (BASELINE configs[1])",
            "stiffness": {"ms": ms_k, "elements_per_s": cells.shape[0] / (ms_k * 1e-3),
                          "roofline": ctx.roof(nb, ms_k, ("k_fused_persist<2,1,sym,lean> (persistent CTAs, block lists prefetched by the bulk-copy engine)" if persist else
                                                              "k_fused_assemble<2,1,sym,lean,nodes> (block-local node copies through the bulk-copy prologue)") if fused else "two-kernel",
                                               bytes_per_element=B_ASM["c2"])},
            "mass": {"ms": ms_m, "elements_per_s": cells.shape[0] / (ms_m * 1e-3),
                     "roofline": ctx.roof(nb, ms_m, ("k_fused_persist<2,1,sym,reac>" if persist else "k_fused_assemble<2,1,sym,reac>") + " (reference tensor R_ij in the constant bank)",
                                          bytes_per_element=B_ASM["c2"])},
            "solve": {"seconds": st["seconds"], "iters": st["iters"], "converged": st["converged"],
                      "rel_resid": st["rel_resid"], "us_per_iter": st["seconds"] / it * 1e6,
                      "roofline": ctx.roof((12 * nnz + 92 * n) * it, st["seconds"] * 1e3, "CG iteration (graph replay)")},
            "parity": {"max_err_vs_exact": float(np.max(np.abs(err))), "max_err_bound": (np.pi / N) ** 2,
                       "solution_ok": bool(np.max(np.abs(err)) < (np.pi / N) ** 2)}}


# ---- C3: 2D advection-diffusion-reaction P2 (non-symmetric), 2 M triangles, BiCGSTAB, 1/2/4/8 GPUs ---------------------
def bench_c3(ctx, args):
    fdb, world, rank = ctx.fdb, ctx.world, ctx.rank
    N = 1000
    pi = np.pi
    nodes, cells, bnd = fdb.meshes.unit_square(N)
    mesh_g = fdb.Triangulation(nodes, cells, bnd)
    basis = fdb.LagrangianBasis(mesh_g, 2)             # global dof table, enumerated on the device
    dofs, nd, bd = basis.dofs(), basis.size(), basis.boundary_dofs()
    L = -fdb.laplacian() + fdb.advection([-1.0, 0.0]) + fdb.reaction(1.0)
    if world > 1:
        loc = fdb.partition.partition_dofs(nodes, cells, dofs, nd, bd, rank, world)
        mesh = fdb.Triangulation(loc.nodes, loc.cells, np.asarray(bnd).ravel()[loc.node_ids])
        nl = loc.n_local_dofs
        s = fdb.Space(mesh, 2, loc.dofs, nl, loc.boundary, pass_cells=True)
        s.set_dof0_rule(loc.owns_dof0)
    else:
        loc, nl = None, nd
        s = fdb.Space(mesh_g, 2, dofs, nd, bd)
    s.set_stream(ctx.stream.cuda_stream)
    s.prepare(False)
    A = fdb.Matrix(s)
    ms_a = ctx.time_loop(lambda: A.assemble(L), 10)
    fused, _ = s.last_path()
    which = s.last_kernel()
    nnz = A.nnz()
    if world > 1:
        A.set_partition(ctx.comm, loc)
        if os.environ.get("FDB_PEER", "1") == "1":
            A.enable_peer_memory(loc, ctx.gather)   # persistent BiCGSTAB: halo pushes + reductions over peer memory
    xy = s.dofs_coords()
    q = s.quadrature_nodes()
    # manufactured: u = sin(pi x) sin(pi y);  L u = 2 pi^2 u - u_x + u
    f = (2 * pi ** 2 + 1) * np.sin(pi * q[:, 0]) * np.sin(pi * q[:, 1]) - pi * np.cos(pi * q[:, 0]) * np.sin(pi * q[:, 1])
    del q
    b, fq = fdb.Vector(nl), fdb.Vector(f.size, f)
    fdb.lib().fdb_assemble_forcing(s.h, fq.h, b.h)
    x = fdb.Vector(nl).fill(0.0)
    A.set_dirichlet(fdb.Vector(nl).fill(0.0), b, x)
    opts = fdb.SolverOptions("bicgstab", rtol=1e-8, maxit=30000, check_every=50)
    ctx.barrier()
    st = A.solve(b, x, opts, raise_on_fail=False)
    t_solve = ctx.max(st["seconds"])
    it = max(st["iters"], 1)
    u_loc = x.download()
    n_own = loc.n_owned if loc is not None else nd
    err = u_loc[:n_own] - (np.sin(pi * xy[:, 0]) * np.sin(pi * xy[:, 1]))[:n_own]
    max_err = ctx.max(float(np.max(np.abs(err))))
    nnz_total = ctx.sum(nnz) if world > 1 else nnz
    nb = B_ASM["c3"] * cells.shape[0]
    bound = 2e-8   # O(h^3) interpolation error of P2 at h = 1e-3, observed 2e-9
    return {"workload": f"2D advection-diffusion-reaction P2, unit square N={N} ({cells.shape[0]} triangles, {nd} dofs), "
                        f"assembly + BiCGSTAB 1e-8 (BASELINE configs[2])",
            "assembly": {"ms": ms_a, "elements_per_s": cells.shape[0] / (ms_a * 1e-3),
                         "roofline": ctx.roof(nb, ms_a, ("k_fused_persist<2,2,nonsym,tensor> (persistent CTAs, block lists prefetched by the bulk-copy engine; compact records, "
                                                         "reference tensors in the constant bank)" if which == 2 else
                                                         "k_fused_assemble<2,2,nonsym,tensor> (compact records, reference tensors in the constant bank)") if fused else
                                              "k_local_assemble<2,2,nonsym,tensor>+k_segmented_reduce<0>", bytes_per_element=B_ASM["c3"])},
            "solve": {"seconds": t_solve, "iters": st["iters"], "converged": st["converged"], "rel_resid": st["rel_resid"],
                      "us_per_iter": t_solve / it * 1e6,
                      "roofline": ctx.roof((24 * nnz_total + 190 * nd) * it, t_solve * 1e3,
                                           "BiCGSTAB iteration: 2 SpMV + fused vector updates / dots" if world == 1 else
                                           "k_bicgstab_sell: one cooperative kernel per rank, halo + reductions over NVLink peer memory")},
            "parity": {"max_err_vs_exact": max_err, "max_err_bound": bound,
                       "solution_ok": bool(st["converged"] and max_err < bound)}}


# ---- C5: 3D reaction-diffusion P2 (extension A10), 20.25 M tets, mass + stiffness assembly, element-partitioned ----------
def bench_c5(ctx, args):
    """Every rank assembles the rows it owns of the 8-way partition (no communication in assembly).  Checked through exact
    identities of the assembled operators (constants in the kernel of every owned stiffness row, row sums of the mass
    matrix = integrals of the basis functions)."""
    fdb, world, rank = ctx.fdb, ctx.world, ctx.rank
    if os.environ.get("FDB_C5_SLAB"):   # development: "rank/world" -> one slab of the partition on a smaller launch
        rank, world = (int(v) for v in os.environ["FDB_C5_SLAB"].split("/"))
    n_cube = 150
    nodes, cells, bnd = fdb.meshes.unit_cube(n_cube)
    mesh = fdb.Triangulation(nodes, cells, bnd)
    basis = fdb.LagrangianBasis(mesh, 2)                      # global edge numbering on the device
    dofs, nd, bd = basis.dofs(), basis.size(), basis.boundary_dofs()
    t0 = time.perf_counter()
    loc = fdb.partition.partition_dofs(nodes, cells, dofs, nd, bd, rank, world)
    t_part = time.perf_counter() - t0
    n_cells_total = cells.shape[0]
    n_nodes_total = nodes.shape[0]
    del basis, dofs, mesh, nodes, cells
    lmesh = fdb.Triangulation(loc.nodes, loc.cells, np.zeros(loc.nodes.shape[0], np.uint8))
    s = fdb.Space(lmesh, 2, loc.dofs, loc.n_local_dofs, loc.boundary, pass_cells=True)
    s.set_stream(ctx.stream.cuda_stream)
    s.prepare(True)
    K, Mm = fdb.Matrix(s), fdb.Matrix(s)
    stiff, mass = -fdb.laplacian(), fdb.reaction(1.0)
    ms_k = ctx.time_loop(lambda: K.assemble(stiff), 5, warm=2)
    fused, launches = s.last_path()
    ms_m = ctx.time_loop(lambda: Mm.assemble(mass), 5, warm=2)
    nl, no = loc.n_local_dofs, loc.n_owned
    ones, y = fdb.Vector(nl).fill(1.0), fdb.Vector(nl)
    K.spmv(ones, y)
    ky = np.abs(y.download()[:no]).max()
    scale = np.abs(K.download_csc()[2]).max()
    Mm.spmv(ones, y)
    my = y.download()[:no]
    cnt = np.bincount(loc.dofs.ravel(), minlength=nl)[:no]
    is_vertex = loc.local_to_global[:no] < n_nodes_total
    vol = (1.0 / n_cube) ** 3 / 6.0
    expect = cnt * vol * np.where(is_vertex, -1.0 / 20.0, 1.0 / 5.0)
    ok = bool(ky < 1e-11 * scale and np.max(np.abs(my - expect)) < 1e-12 * np.abs(expect).max())
    ok = ctx.sum(0.0 if ok else 1.0) == 0.0
    # whole job on `world` GPUs; a development run of one slab on one GPU (FDB_C5_SLAB) earns its share only
    share = ctx.world / world
    nb = B_ASM["c5"] * n_cells_total * share
    kname = ("k_fused_assemble<3,2,sym,{}> (compact records of the needed entries in shared memory, one launch)" if fused
             else "k_local_assemble_p2tet_const<sym,{}> + k_segmented_reduce<sym>")
    return {"workload": f"3D reaction-diffusion P2 (extension A10), unit cube n={n_cube} ({n_cells_total} tets, {nd} dofs), "
                        f"mass + stiffness assembly, element-partitioned over {world} GPUs (BASELINE configs[4])",
            "stiffness": {"ms": ms_k, "elements_per_s": n_cells_total * share / (ms_k * 1e-3),
                          "roofline": ctx.roof(nb, ms_k, kname.format("lap"), bytes_per_element=B_ASM["c5"])},
            "mass": {"ms": ms_m, "elements_per_s": n_cells_total * share / (ms_m * 1e-3),
                     "roofline": ctx.roof(nb, ms_m, kname.format("reac"), bytes_per_element=B_ASM["c5"])},
            "local": {"cells_max": int(ctx.max(loc.cells.shape[0])), "dofs_max": int(ctx.max(nl)),
                      "host_partition_s": ctx.max(t_part)},
            "parity": {"stiffness_rows_annihilate_constants": ok, "mass_row_sums_exact": ok, "solution_ok": ok}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=119, help="cubes per edge (119 -> 10,110,954 tets)")
    ap.add_argument("--ref-n", type=int, default=0, help="reference arm: force this cube size (0 = the workload's n)")
    ap.add_argument("--ref-budget-s", type=float, default=280.0, help="reference arm: wall-clock budget of the run")
    ap.add_argument("--no-ref-solve", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C2 / C3 / C5 blocks")
    ap.add_argument("--min-warmup-s", type=float, default=0.6, help="0 under ncu")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, int(os.environ.get("RANK", "0")))
    args.warmup = max(args.warmup, 3)
    ctx = Ctx(args)
    line = bench_c4(ctx, args)
    if not args.no_extra:
        extra = {}
        want_c5 = ctx.world == 8 or os.environ.get("FDB_BENCH_C5") == "1"
        for name, fn, cond in (("c2", bench_c2, ctx.world == 1), ("c3", bench_c3, True), ("c5", bench_c5, want_c5)):
            if not cond:
                continue
            try:
                extra[name] = fn(ctx, args)
            except Exception as e:  # an extra block must never cost the headline line
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            ctx.torch.cuda.synchronize()
        line["configs"] = extra

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle on 1 core, like the reference ------------------------
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        ns = 40
        sn, sc, sb = ctx.fdb.meshes.unit_cube(ns)
        t0 = time.perf_counter()
        orc.assemble_operator(1, sn, sc, sc, sn.shape[0], [(orc.LAPLACIAN, -1.0)], True)
        t = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sc.shape[0] / t, "unit": "elements/s", "cores": 1, "kind": "port",
                                "host_cores": os.cpu_count(),
                                "sample": f"oracle (-O2 -march=x86-64, 1 thread like the reference) on unit cube "
                                          f"n={ns}: {sc.shape[0]} tets in {t:.1f} s (same mesh family as the workload)"}
    if ctx.rank == 0:
        print(json.dumps(line))
    ctx.comm = None
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
