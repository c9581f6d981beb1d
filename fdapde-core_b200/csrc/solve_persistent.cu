// K8, persistent form: the whole conjugate-gradient loop in ONE cooperative kernel.
//
// Same mathematics and the same deterministic reductions as the multi-kernel loop of solve.cu (which replaces
// FEMLinearEllipticSolver::solve, finite_elements/solvers/fem_linear_elliptic_solver.h:34-50), but the three phases of
// an iteration are separated by grid-wide barriers instead of kernel boundaries:
//     A  q = A p, partial p.q                          | grid barrier
//     B  alpha; x += alpha p; r -= alpha q; z = M^-1 r | grid barrier
//     C  convergence test; beta; p = z + beta p        | grid barrier
// so there is no launch latency or tail effect per phase, scalars never leave the chip, the iteration stops exactly
// at convergence, and -- on several GPUs -- the halo exchange and the dot-product reductions are done by this same
// kernel over NVLink peer memory (cudaIpc-mapped buffers of the neighbouring ranks): phase C pushes the owned
// entries of p that neighbours need straight into their vectors and raises a flag; reductions are an all-gather of
// one cache line per rank into every peer, summed in rank order (identical on all ranks).  No NCCL call, no host
// round trip inside the loop.
#include <cooperative_groups.h>

#include <algorithm>

#include "solve_common.cuh"

namespace cg = cooperative_groups;

namespace fdb {

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ---- flag-in-data exchange over peer memory (see solve_common.cuh) ------------------------------------------------
__device__ __forceinline__ void ll_store(LLWord* dst, double v, unsigned tag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long t = (unsigned long long)tag << 32;
    const unsigned long long lo = (bits & 0xffffffffull) | t, hi = (bits >> 32) | t;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(lo), "l"(hi) : "memory");
}
// spins until both words carry `tag`; *err is raised on timeout
__device__ __forceinline__ double ll_load(const LLWord* src, unsigned tag, int* err) {
    unsigned long long lo, hi, spins = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
        if ((unsigned)(lo >> 32) == tag && (unsigned)(hi >> 32) == tag) break;
        if (++spins > (1ull << 28)) { *err = 1; break; }
    }
    return __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
}

constexpr unsigned long long SPIN_LIMIT = 1ull << 31;
constexpr int PB = 1024;  // threads per block of the persistent kernel: few, fat blocks keep the grid barrier cheap

__device__ __forceinline__ double block_sum_pb(double v, double* sh /* PB/32 doubles */) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0;
#pragma unroll
    for (int k = 0; k < PB / 32; ++k) t += sh[k];
    return t;
}

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// sums three values over the block with one barrier pair; result valid in every thread
__device__ __forceinline__ void block_sum3_pb(double& a, double& b, double& c, double* sh3 /* 3 * PB/32 doubles */) {
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    constexpr int NW = PB / 32;
    __syncthreads();
    if (l == 0) { sh3[w] = a; sh3[NW + w] = b; sh3[2 * NW + w] = c; }
    __syncthreads();
    double ta = 0, tb = 0, tc = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) { ta += sh3[k]; tb += sh3[NW + k]; tc += sh3[2 * NW + k]; }
    a = ta; b = tb; c = tc;
}

// Grid-wide (and, with PEER, cross-rank) sum of up to three values that is ALSO the grid barrier:
//   every block publishes its partials and takes a ticket; the block that takes the last ticket sums all partials
//   in block order (deterministic), exchanges the per-rank sums with the other ranks through their peer-mapped
//   reduction lines (an all-gather of one 128-byte line per rank, summed in rank order, identical on every rank),
//   and publishes the result with a release store of the sequence number; every other block spins on that number
//   with an acquire load.  One atomic and one flag per block: cheaper than a ticket plus a separate grid.sync().
template <bool PEER>
__device__ __forceinline__ void grid_sum3(double& a, double& b, double& c, double* part, unsigned* ticket, int point,
                                          unsigned long long seq, unsigned tag, const PeerView& pv,
                                          double* bcast /* global, [2][4] */,
                                          unsigned long long* bseq /* global */, double* sh3, double* bc) {
    const int np = gridDim.x;
    block_sum3_pb(a, b, c, sh3);
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = a; part[np + blockIdx.x] = b; part[2 * np + blockIdx.x] = c;
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = ((t + 1) % (unsigned)np == 0) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double va = 0, vb = 0, vc = 0;
        for (int k = threadIdx.x; k < np; k += PB) {
            va += __ldcg(part + k); vb += __ldcg(part + np + k); vc += __ldcg(part + 2 * np + k);
        }
        block_sum3_pb(va, vb, vc, sh3);
        if (PEER) {
            // all-gather of this rank's three sums into every rank (flag-in-data stores), then one thread per source
            // rank waits for its triple and the sums are formed in rank order: identical on every rank
            __shared__ double s_red[8][3];
            if ((int)threadIdx.x < pv.world) {
                LLWord* dst = pv.red_of[threadIdx.x] + ((size_t)point * pv.world + pv.rank) * 3;
                ll_store(dst, va, tag); ll_store(dst + 1, vb, tag); ll_store(dst + 2, vc, tag);
                const LLWord* src = pv.my_red + ((size_t)point * pv.world + threadIdx.x) * 3;
                s_red[threadIdx.x][0] = ll_load(src, tag, pv.error);
                s_red[threadIdx.x][1] = ll_load(src + 1, tag, pv.error);
                s_red[threadIdx.x][2] = ll_load(src + 2, tag, pv.error);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                va = 0; vb = 0; vc = 0;
                for (int r = 0; r < pv.world; ++r) { va += s_red[r][0]; vb += s_red[r][1]; vc += s_red[r][2]; }
            }
        }
        if (threadIdx.x == 0) {
            double* o = bcast + (point & 1) * 4;
            o[0] = va; o[1] = vb; o[2] = vc;
            st_release_gpu(bseq, seq);
        }
    }
    if (threadIdx.x == 0) {
        unsigned long long spins = 0;
        while (ld_acquire_gpu(bseq) < seq) {
            if (++spins > SPIN_LIMIT) {   // a block (or rank) never arrived: raise the error word, the host reports it
                if (pv.error) *reinterpret_cast<volatile int*>(pv.error) = 1;
                break;
            }
        }
        const double* o = bcast + (point & 1) * 4;
        bc[0] = __ldcg(o); bc[1] = __ldcg(o + 1); bc[2] = __ldcg(o + 2);
    }
    __syncthreads();
    a = bc[0]; b = bc[1]; c = bc[2];
    __syncthreads();
}

// p is double buffered (p0 / p1 alternate every iteration) so that the new direction and the halo push of the same
// phase never read a value another thread is overwriting.
template <int TPR, bool PEER>
__global__ void __launch_bounds__(PB, 2)
k_cg_persistent(int n, int ld, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const double* __restrict__ val, const double* __restrict__ b, double* __restrict__ x,
                double* __restrict__ r, double* pbuf /* [2][ld] */, double* __restrict__ q, double* __restrict__ z,
                const double* __restrict__ dinv, double* part, unsigned* tickets, double* bcast, Scal* sc,
                double* __restrict__ hist, int hist_cap, int maxit, double rtol, PeerView pv,
                unsigned long long* trace /* nullable: [64 iterations][8 stamps] of %globaltimer */) {
    cg::grid_group grid = cg::this_grid();
#define FDB_STAMP(k) do { if (trace && gtid == 0 && it < 64) trace[it * 8 + (k)] = global_timer_ns(); } while (0)
    __shared__ double sh[3 * (PB / 32)];
    __shared__ double bc[4];
    unsigned long long* bseq = reinterpret_cast<unsigned long long*>(bcast + 8);
    const int np = gridDim.x;
    constexpr int RPB = PB / TPR;
    const int lane = threadIdx.x % TPR, rl = threadIdx.x / TPR;
    const int gtid = blockIdx.x * PB + threadIdx.x, gsz = np * PB;
    unsigned long long seq = 0;        // grid-barrier sequence number (per launch)
    unsigned tag = pv.tag0;            // exchange tag: same on every rank, never reused over the life of the matrix
    unsigned halo_tag = 0;             // tag of the halo values the next SpMV must see
    int halo_buf = 0;                  // which of the two halo receive buffers they arrive in
    int cur = 0;                       // which p buffer holds the current direction
    double* p = pbuf;

    // PEER: store the owned entries neighbours need straight into their halo buffers (flag-in-data: no fence, no flag);
    // each value is recomputed with f(j) so that no other thread's write is read
    auto push_halo = [&](int buf, unsigned t, auto f) {
        for (int i = 0; i < pv.n_nbr; ++i) {
            const int s0 = pv.send_off[i], cnt = pv.send_off[i + 1] - s0;
            LLWord* dst = pv.nbr_halo[i] + (size_t)buf * pv.nbr_n_halo[i];
            for (int k = gtid; k < cnt; k += gsz) ll_store(dst + k, f(__ldg(pv.send_idx + s0 + k)), t);
        }
    };
    // y = A vec over the owned rows.  Columns >= n are halo entries: they are read from the receive buffer, spinning
    // on the value itself until it has arrived -- the exchange overlaps the interior part of the product.
    auto spmv_rows = [&](const double* vec, double* out, double& dot_acc, const double* w) {
        const LLWord* halo = PEER ? pv.my_halo + (size_t)halo_buf * pv.n_halo : nullptr;
        for (int base = blockIdx.x * RPB; base < n; base += np * RPB) {
            const int row = base + rl;
            double s = 0;
            if (row < n) {
                const int t0 = rowptr[row], t1 = rowptr[row + 1];
                // plain (L1-cached) gathers: the grid barrier before this phase invalidates L1, and the 60 % L1 hit
                // rate of the x gathers is what keeps SpMV at its HBM rate
                for (int t = t0 + lane; t < t1; t += TPR) {
                    const int c = __ldg(colidx + t);
                    double xv;
                    if (PEER && c >= n) xv = ll_load(halo + (c - n), halo_tag, pv.error);
                    else xv = vec[c];
                    s += __ldg(val + t) * xv;
                }
            }
#pragma unroll
            for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (row < n && lane == 0) {
                out[row] = s;
                if (w) dot_acc += s * __ldcg(w + row);
            }
        }
    };

    // ---- initial residual: q = A x ; r = b - q ; z = M^-1 r ; p = z -------------------------------------------------
    if (PEER) {
        ++tag;
        push_halo(0, tag, [&](int j) { return x[j]; });
        halo_tag = tag; halo_buf = 0;
    }
    {
        double dummy = 0;
        spmv_rows(x, q, dummy, nullptr);
    }
    grid.sync();
    double rz = 0, rr = 0, bb = 0;
    double* p1 = pbuf + ld;
    for (int i = gtid; i < n; i += gsz) {
        const double bi = b[i], ri = bi - __ldcg(q + i);
        const double zi = dinv ? dinv[i] * ri : ri;
        r[i] = ri;
        if (dinv) z[i] = zi;
        p1[i] = zi;
        rz += ri * zi; rr += ri * ri; bb += bi * bi;
    }
    if (PEER) {
        ++tag;
        push_halo(1, tag, [&](int j) { const double ri = b[j] - __ldcg(q + j); return dinv ? dinv[j] * ri : ri; });
        halo_tag = tag; halo_buf = 1;
    }
    cur = 1;
    p = p1;
    ++seq; ++tag;
    grid_sum3<PEER>(rz, rr, bb, part, tickets, (int)(tag & 3), seq, tag, pv, bcast, bseq, sh, bc);
    double rz_old = rz;
    const double thr = rtol * rtol * bb;
    int it = 0;
    bool conv = (rr <= thr) || (bb == 0.0), bad = !isfinite(rr) || !isfinite(bb);
    const double* zz = dinv ? z : r;

    while (!conv && !bad && it < maxit) {
        // ---- A: q = A p, p.q ----------------------------------------------------------------------------------------
        double pq = 0, d1 = 0, d2 = 0;
        FDB_STAMP(0);
        spmv_rows(p, q, pq, p);
        FDB_STAMP(1);
        ++seq; ++tag;
        grid_sum3<PEER>(pq, d1, d2, part, tickets, (int)(tag & 3), seq, tag, pv, bcast, bseq, sh, bc);
        FDB_STAMP(2);
        // ---- B: x, r, z -----------------------------------------------------------------------------------------------
        const double alpha = rz_old / pq;
        double rz_new = 0;
        rr = 0;
        d2 = 0;
        for (int i = gtid; i < n; i += gsz) {
            x[i] += alpha * __ldcg(p + i);
            const double ri = r[i] - alpha * __ldcg(q + i);
            r[i] = ri;
            const double zi = dinv ? dinv[i] * ri : ri;
            if (dinv) z[i] = zi;
            rz_new += ri * zi;
            rr += ri * ri;
        }
        FDB_STAMP(3);
        ++seq; ++tag;
        grid_sum3<PEER>(rz_new, rr, d2, part, tickets, (int)(tag & 3), seq, tag, pv, bcast, bseq, sh, bc);
        FDB_STAMP(4);
        // ---- C: convergence, new direction (into the other p buffer) + halo push --------------------------------------
        if (gtid == 0) hist[it % hist_cap] = rr;
        conv = rr <= thr;
        if (!isfinite(rr)) { ++it; bad = true; break; }   // same sums on every thread of every rank: a uniform exit
        if (!conv) {
            const double beta = rz_new / rz_old;
            double* pn = pbuf + (size_t)(cur ^ 1) * ld;
            for (int i = gtid; i < n; i += gsz) pn[i] = __ldcg(zz + i) + beta * __ldcg(p + i);
            if (PEER) {
                ++tag;
                push_halo(cur ^ 1, tag, [&](int j) { return __ldcg(zz + j) + beta * __ldcg(p + j); });
                halo_tag = tag; halo_buf = cur ^ 1;
            }
            rz_old = rz_new;
            cur ^= 1;
            p = pn;
            FDB_STAMP(5);
            grid.sync();
            FDB_STAMP(6);
            FDB_STAMP(7);
        }
        ++it;
    }
#undef FDB_STAMP
    if (gtid == 0) {
        sc->pad = (int)tag;  // last tag used: the host carries it to the next solve
        sc->bb = bb; sc->thr = thr; sc->rr = rr; sc->iters = it;
        sc->done = conv ? 1 : 0;
        sc->breakdown = (bb == 0.0) ? 2 : (bad ? 3 : 0);
    }
}

// ---- single-reduction CG (Chronopoulos & Gear 1989), opt-in with FDB_CG1=1 ------------------------------------------
// The standard recurrence needs two dependent reductions per iteration (p.Ap, then r.z); across GPUs each one is an
// all-gather over NVLink plus a grid barrier, and at 1.7 M unknowns on 8 GPUs those latencies are most of the iteration.
// With w = A z the same Krylov iterates follow from ONE fused reduction of (r.z, w.z, r.r):
//     beta = gamma / gamma_old,   alpha = gamma / (delta - beta gamma / alpha_old)
//     p = z + beta p,  s = w + beta s,  x += alpha p,  r -= alpha s,  z = M^-1 r
// (s tracks A p without a second product).  Per iteration: one SpMV, one cross-GPU reduction, one grid barrier.
// r and s are double buffered so that the halo push can recompute a neighbour's z_j from values no thread overwrites.
template <int TPR, bool PEER>
__global__ void __launch_bounds__(PB, 2)
k_cg1_persistent(int n, int ld, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                 const double* __restrict__ val, const double* __restrict__ b, double* __restrict__ x,
                 double* rbuf /* [2][ld] */, double* sbuf /* [2][ld] */, double* zbuf /* [2][ld], Jacobi only */,
                 double* __restrict__ p, double* __restrict__ w, const double* __restrict__ dinv, double* part,
                 unsigned* tickets, double* bcast, Scal* sc, double* __restrict__ hist, int hist_cap, int maxit,
                 double rtol, PeerView pv) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[3 * (PB / 32)];
    __shared__ double bc[4];
    unsigned long long* bseq = reinterpret_cast<unsigned long long*>(bcast + 8);
    const int np = gridDim.x;
    constexpr int RPB = PB / TPR;
    const int lane = threadIdx.x % TPR, rl = threadIdx.x / TPR;
    const int gtid = blockIdx.x * PB + threadIdx.x, gsz = np * PB;
    unsigned long long seq = 0;
    unsigned tag = pv.tag0;
    unsigned halo_tag = 0;
    int halo_buf = 0;

    auto push_halo = [&](int buf, unsigned t, auto f) {
        for (int i = 0; i < pv.n_nbr; ++i) {
            const int s0 = pv.send_off[i], cnt = pv.send_off[i + 1] - s0;
            LLWord* dst = pv.nbr_halo[i] + (size_t)buf * pv.nbr_n_halo[i];
            for (int k = gtid; k < cnt; k += gsz) ll_store(dst + k, f(__ldg(pv.send_idx + s0 + k)), t);
        }
    };
    // out = A vec over the owned rows, dot_acc += out . vec (lane 0 of every row)
    auto spmv_rows = [&](const double* vec, double* out, double& dot_acc, bool with_dot) {
        const LLWord* halo = PEER ? pv.my_halo + (size_t)halo_buf * pv.n_halo : nullptr;
        for (int base = blockIdx.x * RPB; base < n; base += np * RPB) {
            const int row = base + rl;
            double s = 0;
            if (row < n) {
                const int t0 = rowptr[row], t1 = rowptr[row + 1];
                for (int t = t0 + lane; t < t1; t += TPR) {
                    const int c = __ldg(colidx + t);
                    double xv;
                    if (PEER && c >= n) xv = ll_load(halo + (c - n), halo_tag, pv.error);
                    else xv = vec[c];
                    s += __ldg(val + t) * xv;
                }
            }
#pragma unroll
            for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (row < n && lane == 0) {
                out[row] = s;
                if (with_dot) dot_acc += s * vec[row];
            }
        }
    };

    // ---- r0 = b - A x0, z0 = M^-1 r0, p = s = 0 -------------------------------------------------------------------------
    if (PEER) {
        ++tag;
        push_halo(0, tag, [&](int j) { return x[j]; });
        halo_tag = tag; halo_buf = 0;
    }
    {
        double dummy = 0;
        spmv_rows(x, w, dummy, false);
    }
    grid.sync();
    int cur = 0;                                   // which r / s / z buffers hold the current iterate
    double* r = rbuf;
    double* z = dinv ? zbuf : rbuf;
    double gam = 0, rr = 0, bb = 0;                // per-thread partials of r.z, r.r (carried to the next reduction)
    for (int i = gtid; i < n; i += gsz) {
        const double bi = b[i], ri = bi - __ldcg(w + i);
        const double zi = dinv ? dinv[i] * ri : ri;
        r[i] = ri;
        if (dinv) z[i] = zi;
        p[i] = 0.0;
        sbuf[i] = 0.0;
        gam += ri * zi; rr += ri * ri; bb += bi * bi;
    }
    if (PEER) {
        ++tag;
        push_halo(1, tag, [&](int j) { const double ri = b[j] - __ldcg(w + j); return dinv ? dinv[j] * ri : ri; });
        halo_tag = tag; halo_buf = 1;
    }
    double gam_p = gam, rr_p = rr;                 // keep the partials: the loop's first reduction needs them again
    ++seq; ++tag;
    grid_sum3<PEER>(gam, rr, bb, part, tickets, (int)(tag & 3), seq, tag, pv, bcast, bseq, sh, bc);
    const double thr = rtol * rtol * bb;
    int it = 0;
    bool conv = (rr <= thr) || (bb == 0.0), bad = !isfinite(rr) || !isfinite(bb);
    double gam_old = 1.0, alpha = 1.0;

    while (!conv && !bad && it < maxit) {
        // ---- A: w = A z, then ONE reduction of (r.z, w.z, r.r) ------------------------------------------------------------
        double del = 0;
        spmv_rows(z, w, del, true);
        gam = gam_p; rr = rr_p;
        ++seq; ++tag;
        grid_sum3<PEER>(gam, del, rr, part, tickets, (int)(tag & 3), seq, tag, pv, bcast, bseq, sh, bc);
        if (!isfinite(rr) || !isfinite(del)) { bad = true; break; }   // same sums everywhere: a uniform exit
        if (it > 0 && gtid == 0) hist[(it - 1) % hist_cap] = rr;
        conv = rr <= thr;
        if (conv) break;
        const double beta = it == 0 ? 0.0 : gam / gam_old;
        alpha = it == 0 ? gam / del : gam / (del - beta * gam / alpha);
        gam_old = gam;
        // ---- C: p, s, x, r, z; partials of the next reduction; halo push of the new z -----------------------------------
        const double* r_old = rbuf + (size_t)cur * ld;
        const double* s_old = sbuf + (size_t)cur * ld;
        double* r_new = rbuf + (size_t)(cur ^ 1) * ld;
        double* s_new = sbuf + (size_t)(cur ^ 1) * ld;
        double* z_new = dinv ? zbuf + (size_t)(cur ^ 1) * ld : r_new;
        gam_p = 0; rr_p = 0;
        for (int i = gtid; i < n; i += gsz) {
            const double zi = z[i];
            const double pi = zi + beta * p[i];
            const double si = __ldcg(w + i) + beta * s_old[i];
            p[i] = pi;
            s_new[i] = si;
            x[i] += alpha * pi;
            const double ri = r_old[i] - alpha * si;
            r_new[i] = ri;
            const double zn = dinv ? dinv[i] * ri : ri;
            if (dinv) z_new[i] = zn;
            gam_p += ri * zn;
            rr_p += ri * ri;
        }
        if (PEER) {
            ++tag;
            push_halo(halo_buf ^ 1, tag, [&](int j) {
                const double ri = r_old[j] - alpha * (__ldcg(w + j) + beta * s_old[j]);
                return dinv ? dinv[j] * ri : ri;
            });
            halo_tag = tag; halo_buf ^= 1;
        }
        cur ^= 1;
        r = r_new;
        z = z_new;
        ++it;
        grid.sync();
    }
    if (!conv && !bad && it >= maxit) {   // budget exhausted: the residual of the last update has not been reduced yet
        gam = gam_p; rr = rr_p;
        double d3 = 0;
        ++seq; ++tag;
        grid_sum3<PEER>(gam, d3, rr, part, tickets, (int)(tag & 3), seq, tag, pv, bcast, bseq, sh, bc);
        conv = rr <= thr;
    }
    if (gtid == 0) {
        sc->pad = (int)tag;
        sc->bb = bb; sc->thr = thr; sc->rr = rr; sc->iters = it;
        sc->done = conv ? 1 : 0;
        sc->breakdown = (bb == 0.0) ? 2 : (bad ? 3 : 0);
    }
}

// =====================================================================================================================
template <int TPR, bool PEER>
static int launch_persistent(fdb_matrix* A, int grid, int n, const double* b, double* x, double* W, size_t ld,
                             const double* dv, double* part, unsigned* tickets, double* bcast, Scal* sc, double* hist,
                             int hist_cap, int maxit, double rtol, const PeerView& pv, bool cg1) {
    fdb_space* s = A->space;
    const Pattern* P = A->pat;
    if (cg1) {  // single-reduction variant: W = [r0 r1 | s0 s1 | z0 z1 | p | w | dinv]
        double *rbuf = W, *sbuf = W + 2 * ld, *zbuf = W + 4 * ld, *pp = W + 6 * ld, *ww = W + 7 * ld;
        const int32_t* rowptr1 = P->rowptr.p;
        const int32_t* colidx1 = P->colidx.p;
        const double* val1 = A->val.p;
        int ldi1 = (int)ld;
        PeerView pv1 = pv;
        void* args1[] = {&n, &ldi1, &rowptr1, &colidx1, &val1, &b, &x, &rbuf, &sbuf, &zbuf, &pp, &ww, &dv, &part, &tickets,
                         &bcast, &sc, &hist, &hist_cap, &maxit, &rtol, &pv1};
        FDB_CUDA(cudaLaunchCooperativeKernel((void*)k_cg1_persistent<TPR, PEER>, dim3(grid), dim3(PB), args1, 0, s->stream));
        return FDB_OK;
    }
    // W = [r | p0 | p1 | q | z]; with PEER the two p buffers are the peer-visible part
    double *r = W, *pbuf = W + ld, *q = W + 3 * ld, *z = W + 4 * ld;
    const int32_t* rowptr = P->rowptr.p;
    const int32_t* colidx = P->colidx.p;
    const double* val = A->val.p;
    int ldi = (int)ld;
    PeerView pvc = pv;
    static DevBuf<unsigned long long> trace_buf;
    unsigned long long* trace = nullptr;
    const bool tracing = getenv("FDB_CG_TRACE") != nullptr;
    if (tracing) {
        if (!trace_buf.p) FDB_TRY(trace_buf.alloc(64 * 8));
        FDB_CUDA(cudaMemsetAsync(trace_buf.p, 0, 64 * 8 * 8, s->stream));
        trace = trace_buf.p;
    }
    void* args[] = {&n, &ldi, &rowptr, &colidx, &val, &b, &x, &r, &pbuf, &q, &z, &dv, &part, &tickets, &bcast, &sc, &hist,
                    &hist_cap, &maxit, &rtol, &pvc, &trace};
    FDB_CUDA(cudaLaunchCooperativeKernel((void*)k_cg_persistent<TPR, PEER>, dim3(grid), dim3(PB), args, 0, s->stream));
    if (tracing) {  // debug aid: average phase durations (ns) over iterations 8..63, printed by rank-local stderr
        std::vector<unsigned long long> h(64 * 8);
        FDB_CUDA(cudaStreamSynchronize(s->stream));
        FDB_CUDA(cudaMemcpy(h.data(), trace_buf.p, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
        double acc[8] = {0};
        int cnt = 0;
        for (int it = 8; it < 63; ++it) {
            if (!h[it * 8 + 7] || !h[(it + 1) * 8]) break;
            for (int k = 0; k < 7; ++k) acc[k] += (double)(h[it * 8 + k + 1] - h[it * 8 + k]);
            acc[7] += (double)(h[(it + 1) * 8] - h[it * 8 + 7]);
            ++cnt;
        }
        if (cnt)
            fprintf(stderr, "[fdb] CG trace (us, %d iterations, block 0): spmv %.1f | reduce(pq) %.1f | update %.1f | reduce(rz,rr) %.1f | "
                    "direction+push %.1f | barrier %.1f | halo wait %.1f | loop %.1f\n", cnt, acc[0] / cnt / 1e3, acc[1] / cnt / 1e3,
                    acc[2] / cnt / 1e3, acc[3] / cnt / 1e3, acc[4] / cnt / 1e3, acc[5] / cnt / 1e3, acc[6] / cnt / 1e3, acc[7] / cnt / 1e3);
    }
    return FDB_OK;
}

template <int TPR, bool PEER> static int max_grid(const fdb_space* s, int* grid, bool cg1) {
    int per_sm = 0;
    if (cg1) FDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg1_persistent<TPR, PEER>, PB, 0));
    else FDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_persistent<TPR, PEER>, PB, 0));
    if (const char* e = getenv("FDB_PERSISTENT_BPS")) per_sm = std::min(per_sm, std::max(1, atoi(e)));  // experiment
    *grid = per_sm * s->sm_count;
    return FDB_OK;
}

__global__ void k_jacobi_p(int n, const int32_t* __restrict__ diag, const double* __restrict__ val,
                           double* __restrict__ dinv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = diag[i];
    double a = d >= 0 ? val[d] : 0.0;
    dinv[i] = a != 0.0 ? 1.0 / a : 1.0;
}

int& persistent_mode() {
    static int mode = -1;  // -1: from the environment, 0: never, 1: multi-GPU only (default), 2: always
    return mode;
}

int solve_cg_persistent(fdb_matrix* A, const double* b, double* x, const fdb_solver_opts* o, fdb_solve_stats* stats,
                        bool* handled) {
    *handled = false;
    fdb_space* s = A->space;
    Partition* part = A->part;
    // One GPU: the multi-kernel loop is (slightly) faster -- 103 vs 107 us/iteration at C4: the grid barriers cost more than
    // the launches they replace -- so the persistent kernel is opt-in there (FDB_PERSISTENT=1).  Several GPUs: it is
    // the default as soon as the peer-memory plan exists, because it removes every NCCL call from the loop.
    int& mode = persistent_mode();
    if (mode < 0) mode = getenv("FDB_NO_PERSISTENT") ? 0 : (getenv("FDB_PERSISTENT") ? 2 : 1);
    if (mode == 0) return FDB_OK;
    if (!part && mode != 2) return FDB_OK;
    if (part && !part->peer_ready) return FDB_OK;  // partitioned without peer memory: NCCL multi-kernel loop
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s->device);
    if (!coop) return FDB_OK;
    const Pattern* P = A->pat;
    cudaStream_t st = s->stream;
    const size_t ld = (size_t)s->n_dofs;
    const int n = part ? part->n_owned : s->n_dofs;
    const long long n_glob = part ? part->n_global : (long long)s->n_dofs;   // identical on every rank
    const int maxit = o->maxit > 0 ? o->maxit : (int)std::min<long long>(10 * std::max<long long>(n_glob, 1), 2000000000LL);
    const bool jac = o->jacobi != 0;
    const int hist_cap = 1 << 16;
    const int tpr = pick_tpr(P, n);
    const bool peer = part != nullptr;
    const bool cg1 = getenv("FDB_CG1") != nullptr;   // single-reduction recurrence (k_cg1_persistent), opt-in
    int grid = 0;
#define FDB_GRID(T) FDB_TRY((peer ? max_grid<T, true>(s, &grid, cg1) : max_grid<T, false>(s, &grid, cg1)))
    switch (tpr) {
    case 1: FDB_GRID(1); break;
    case 2: FDB_GRID(2); break;
    case 4: FDB_GRID(4); break;
    case 8: FDB_GRID(8); break;
    case 16: FDB_GRID(16); break;
    default: FDB_GRID(32); break;
    }
#undef FDB_GRID
    if (grid <= 0) return FDB_OK;
    if (A->work.n < 9 * ld) FDB_TRY(A->work.alloc(9 * ld));
    if (A->partials.n < 8 * (size_t)grid + 128) FDB_TRY(A->partials.alloc(8 * (size_t)grid + 128));
    if (A->hist.n < (size_t)hist_cap) FDB_TRY(A->hist.alloc((size_t)hist_cap));
    double* W = A->work.p;
    double* dinv = A->work.p + (cg1 ? 8 : 5) * ld;
    double* PA = A->partials.p;
    Scal* sc = reinterpret_cast<Scal*>(PA + 8 * (size_t)grid);
    double* bcast = PA + 8 * (size_t)grid + 16;                                   // [2][4] broadcast slots
    unsigned* tickets = reinterpret_cast<unsigned*>(PA + 8 * (size_t)grid + 32);  // one arrival counter
    FDB_CUDA(cudaMemsetAsync(bcast, 0, 128 + 64, st));  // broadcast slots + their sequence number + the ticket counter
    if (jac) {
        k_jacobi_p<<<(n + 255) / 256, 256, 0, st>>>(n, P->diag.p, A->val.p, dinv);
        FDB_CUDA(cudaGetLastError());
    }
    const double* dv = jac ? dinv : nullptr;
    PeerView pv;
    memset(&pv, 0, sizeof(pv));
    if (peer) {
        pv = *reinterpret_cast<PeerView*>(part->peer_view);
        pv.tag0 = part->peer_tag;  // every rank runs the same sequence of solves on this matrix
    }
    cudaEvent_t ev0, ev1;
    FDB_CUDA(cudaEventCreate(&ev0));
    FDB_CUDA(cudaEventCreate(&ev1));
    FDB_CUDA(cudaEventRecord(ev0, st));
    int rc = FDB_OK;
#define FDB_RUN(T)                                                                                                       \
    rc = peer ? launch_persistent<T, true>(A, grid, n, b, x, W, ld, dv, PA, tickets, bcast, sc, A->hist.p, hist_cap, maxit, \
                                           o->rtol, pv, cg1)                                                              \
              : launch_persistent<T, false>(A, grid, n, b, x, W, ld, dv, PA, tickets, bcast, sc, A->hist.p, hist_cap,     \
                                            maxit, o->rtol, pv, cg1)
    switch (tpr) {
    case 1: FDB_RUN(1); break;
    case 2: FDB_RUN(2); break;
    case 4: FDB_RUN(4); break;
    case 8: FDB_RUN(8); break;
    case 16: FDB_RUN(16); break;
    default: FDB_RUN(32); break;
    }
#undef FDB_RUN
    if (rc != FDB_OK) return rc;
    FDB_CUDA(cudaEventRecord(ev1, st));
    Scal h;
    FDB_CUDA(cudaMemcpyAsync(&h, sc, sizeof(Scal), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (peer) {
        part->peer_tag = (unsigned)h.pad;
        int err = 0;
        FDB_CUDA(cudaMemcpy(&err, pv.error, sizeof(int), cudaMemcpyDeviceToHost));
        if (err) FDB_CUDA(cudaMemset(pv.error, 0, sizeof(int)));   // the word is sticky inside a launch, not across solves
        FDB_CHECK(err == 0, FDB_ERR_CUDA, "peer-memory wait timed out (a neighbouring rank did not arrive)");
    }
    if (h.breakdown == 2) {
        FDB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * n, st));
        FDB_CUDA(cudaStreamSynchronize(st));
        h.rr = 0;
    }
    *handled = true;
    const bool converged = h.rr <= h.thr;
    if (stats) {
        stats->iters = h.iters;
        stats->converged = converged ? 1 : 0;
        stats->rel_resid = h.bb > 0 ? sqrt(h.rr / h.bb) : 0.0;
        stats->seconds = ms * 1e-3;
    }
    if (!converged) {
        set_error("iterative solver did not reach the requested tolerance");
        return FDB_ERR_NOT_CONVERGED;
    }
    return FDB_OK;
}

}  // namespace fdb
