// K8, persistent sliced-ELL form: whole Krylov loops (single-reduction CG, BiCGSTAB) in ONE cooperative kernel per GPU.
//
// Replaces the solve of FEMLinearEllipticSolver::solve (finite_elements/solvers/fem_linear_elliptic_solver.h:34-50) on one
// or several GPUs.  What distinguishes it from the first persistent kernel (solve_persistent.cu):
//   * SpMV runs on the sliced-ELL view of the OWNED rows (one thread per row, fully coalesced value / column streams),
//     also for partitioned matrices: halo columns (>= n_owned) are read from the peer-written receive buffer, spinning
//     on the value itself (flag-in-data), and the slices that touch the halo are visited LAST so that the exchange
//     overlaps the interior rows;
//   * one block per SM and a reduction that IS the grid barrier: every block publishes its partial sums and takes a
//     ticket, the last block of each rank sums them in block order and stores the rank's sums straight into every
//     rank's reduction line (all-gather over NVLink, flag-in-data); every block of every rank then reads the `world`
//     lines from its own memory and adds them in rank order -- identical bits on every rank, no broadcast hop, no NCCL
//     call, no host round trip.  A plain barrier is the same reduction with nothing to add;
//   * CG uses the single-reduction recurrence (Chronopoulos & Gear): one SpMV, ONE reduction and one plain barrier per
//     iteration; BiCGSTAB needs three reductions and two plain barriers.
// The recurrences, breakdown tests and stopping rule are those of the multi-kernel loops in solve.cu.
#include <cooperative_groups.h>

#include <algorithm>

#include "solve_common.cuh"

namespace fdb {

constexpr int PT = 1024;   // threads per block; one block per SM keeps the arrival count of a barrier at #SMs
constexpr int NR = 4;      // values per reduction (unused ones are zero)

__device__ __forceinline__ void llp_store(LLWord* dst, double v, unsigned tag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long t = (unsigned long long)tag << 32;
    const unsigned long long lo = (bits & 0xffffffffull) | t, hi = (bits >> 32) | t;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(lo), "l"(hi) : "memory");
}
// spins until both words carry `tag`; *err is raised on timeout (the caller's loop then ends at the next reduction)
__device__ __forceinline__ double llp_load(const LLWord* src, unsigned tag, int* err) {
    unsigned long long lo, hi, spins = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
        if ((unsigned)(lo >> 32) == tag && (unsigned)(hi >> 32) == tag) break;
        ++spins;
        if ((spins & 1023u) == 0) {   // the error word is sticky: once any wait has timed out, every other wait gives up quickly
            if (*reinterpret_cast<volatile int*>(err)) break;
            if (spins > (1ull << 24)) { *reinterpret_cast<volatile int*>(err) = 1; break; }
        }
    }
    return __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
}

struct SellDev {
    const int32_t* ptr;     // slice offsets (slots)
    const void* cols;       // int32 columns, or int16 offsets from the row (C16)
    const double* val;      // values in slot order
    const int32_t* order;   // visiting order of the owned slices (interior first); nullptr = identity
    int n_slices;           // slices covering the owned rows
};

struct RedDev {
    double* part;           // [NR][gridDim.x] per-block partials
    unsigned* ticket;       // arrival counter of the reductions (monotone over the launch)
    unsigned* ticket_b;     // arrival counter of the rank-local barriers
    unsigned* flag_b;       // number of completed rank-local barriers
};

// Rank-local grid barrier with release / acquire semantics (L1 invalidated): orders this rank's own vector updates
// before its next SpMV.  Nothing crosses NVLink here -- remote data (halo entries) synchronises itself through its tags.
__device__ __forceinline__ void grid_barrier_local(const RedDev& R, unsigned& nbar) {
    ++nbar;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t;
        asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(t) : "l"(R.ticket_b) : "memory");
        if ((t + 1) % gridDim.x == 0) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(R.flag_b), "r"(nbar) : "memory");
        } else {
            unsigned v, spins = 0;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(R.flag_b) : "memory");
            } while ((int)(v - nbar) < 0 && ++spins < (1u << 30));
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

// Sum of NR values over all threads of all blocks of all ranks; also a grid-wide (and cross-rank) barrier with acquire
// semantics: data written by any thread before the call is visible to every thread after it (L1 is invalidated).
// Returns the (sticky) error word: non-zero once a wait on a peer has timed out -- the same value in every thread of a block.
// Every sum has a fixed shape (warp tree, then a tree over the warps, then block order, then rank order): bit-reproducible.
template <bool PEER>
__device__ __forceinline__ int grid_sum4(double (&v)[NR], const RedDev& R, unsigned tag, const PeerView& pv,
                                         double* sh /* NR * 32 */, double (*s_red)[NR] /* [9][NR] */, int* s_last) {
    const int np = gridDim.x, tid = threadIdx.x, w = tid >> 5, l = tid & 31;
#pragma unroll
    for (int c = 0; c < NR; ++c) v[c] = warp_sum(v[c]);
    __syncthreads();
    if (l == 0) {
#pragma unroll
        for (int c = 0; c < NR; ++c) sh[c * 32 + w] = v[c];
    }
    __syncthreads();
    if (w == 0) {   // warp 0: tree over the 32 warp sums of each value, then one release-ordered ticket
        double t[NR];
#pragma unroll
        for (int c = 0; c < NR; ++c) t[c] = warp_sum(sh[c * 32 + l]);
        if (l == 0) {
#pragma unroll
            for (int c = 0; c < NR; ++c) R.part[c * np + blockIdx.x] = t[c];
            unsigned tk;   // release: the block's writes (made visible to this thread by the barrier above) precede the ticket
            asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(tk) : "l"(R.ticket) : "memory");
            *s_last = ((tk + 1) % (unsigned)np == 0) ? 1 : 0;
        }
    }
    __syncthreads();
    const int point = (int)(tag & 3u);
    if (*s_last && w < NR) {   // the block that arrived last: warp c sums the partials of value c in block order ...
        if (l == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        __syncwarp();
        double a = 0;
        for (int k = l; k < np; k += 32) a += __ldcg(R.part + w * np + k);
        a = warp_sum(a);
        // ... and lane r stores the rank's sum straight into rank r's line (all-gather over NVLink, flag-in-data)
        if (l < pv.world) llp_store(pv.red_of[l] + ((size_t)point * pv.world + pv.rank) * NR + w, a, tag);
    }
    if (tid < pv.world * NR) {  // every block: thread (r, c) waits for value c of rank r in this rank's own memory
        const int r = tid / NR, c = tid % NR;
        s_red[r][c] = llp_load(pv.my_red + ((size_t)point * pv.world + r) * NR + c, tag, pv.error);
    }
    if (w == 0) {
        __syncwarp();
        if (l == 0) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");   // acquire: also drops stale L1 lines of this SM (CCTL.IVALL)
            s_red[8][0] = (double)*reinterpret_cast<volatile int*>(pv.error);
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NR; ++c) {
        double t = 0;
        for (int r = 0; r < pv.world; ++r) t += s_red[r][c];
        v[c] = t;
    }
    const int err = s_red[8][0] != 0.0;
    __syncthreads();
    return err;
}

// PEER: store the owned entries neighbours need straight into their halo buffers (flag-in-data: no fence, no flag);
// each value is recomputed with f(j), so no other thread's write is read
template <class F>
__device__ __forceinline__ void push_halo(const PeerView& pv, int buf, unsigned tag, int gtid, int gsz, F f) {
    for (int i = 0; i < pv.n_nbr; ++i) {
        const int s0 = pv.send_off[i], cnt = pv.send_off[i + 1] - s0;
        LLWord* dst = pv.nbr_halo[i] + (size_t)buf * pv.nbr_n_halo[i];
        for (int k = gtid; k < cnt; k += gsz) llp_store(dst + k, f(__ldg(pv.send_idx + s0 + k)), tag);
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// out = A vec over the owned rows, one thread per row of a 32-row slice, one warp per slice.  acc(row, value) collects
// the fused dot products.  Columns >= n are halo entries, read from the receive buffer with their arrival tag.
// The kernel is latency bound at the sizes a rank holds on 8 GPUs, so the loop is built for memory-level parallelism:
// eight (value, column) pairs are requested at once, then their eight x entries, then the FMAs (in slot order: the same
// sum as a sequential loop); the header of the warp's NEXT slice is loaded one slice ahead and its value / column lines
// are prefetched into L2 while the current slice is processed.
template <bool PEER, bool C16, class Acc>
__device__ __forceinline__ void spmv_sell(const SellDev& S, int n, const double* vec, double* __restrict__ out,
                                          const LLWord* halo, unsigned halo_tag, int* err, Acc acc) {
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * PT) >> 5;
    int k = (blockIdx.x * PT + threadIdx.x) >> 5;
    if (k >= S.n_slices) return;
    const unsigned long long pol = l2_evict_first_policy();   // matrix streams: read once per SpMV
    int sl = S.order ? __ldg(S.order + k) : k;
    int base = __ldg(S.ptr + sl), len = (__ldg(S.ptr + sl + 1) - base) >> 5;
    for (;;) {
        const int kn = k + nwarps;
        int sln = 0, basen = 0, lenn = 0;
        if (kn < S.n_slices) {
            sln = S.order ? __ldg(S.order + kn) : kn;
            basen = __ldg(S.ptr + sln);
            lenn = (__ldg(S.ptr + sln + 1) - basen) >> 5;
            const char* vp = reinterpret_cast<const char*>(S.val + basen);                       // lenn * 256 bytes
            for (int q = lane; q < 2 * lenn; q += 32) prefetch_l2(vp + 128 * q);
            const char* cp = static_cast<const char*>(S.cols) + (size_t)basen * (C16 ? 2 : 4);   // lenn * 64 / 128 bytes
            const int ncl = C16 ? (lenn + 1) >> 1 : lenn;
            for (int q = lane; q < ncl; q += 32) prefetch_l2(cp + 128 * q);
        }
        const int row = sl * 32 + lane;
        if (row < n) {   // false only in the last, partial slice
            double sum = 0;
            for (int j = 0; j < len; j += 8) {
                double a[8], xv[8];
                int c[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool live = j + u < len;
                    const int slot = base + 32 * (j + u) + lane;
                    a[u] = live ? ld_stream(S.val + slot, pol) : 0.0;
                    if constexpr (C16) c[u] = live ? row + ld_stream(static_cast<const int16_t*>(S.cols) + slot, pol) : row;
                    else c[u] = live ? ld_stream(static_cast<const int32_t*>(S.cols) + slot, pol) : row;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (PEER && c[u] >= n) xv[u] = llp_load(halo + (c[u] - n), halo_tag, err);
                    else xv[u] = vec[c[u]];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) sum += a[u] * xv[u];
            }
            out[row] = sum;
            acc(row, sum);
        }
        if (kn >= S.n_slices) break;
        k = kn; sl = sln; base = basen; len = lenn;
    }
}

struct KrylovArgs {
    int n, ld;                 // owned rows, leading dimension of the workspace vectors
    const double* b;
    double* x;
    double* W;                 // workspace vectors [.][ld]
    const double* dinv;        // Jacobi: 1 / diagonal, or nullptr
    Scal* sc;
    double* hist;
    int hist_cap, maxit;
    double rtol;
    unsigned long long* trace;   // nullable debug aid (FDB_CG_TRACE=1): [64 iterations][8] %globaltimer stamps of block 0
};
__device__ __forceinline__ unsigned long long timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define FDB_STAMP(k) do { if (K.trace && gtid == 0 && it < 64) K.trace[it * 8 + (k)] = timer_ns(); } while (0)

// ---- single-reduction CG (Chronopoulos & Gear 1989) ---------------------------------------------------------------------
// With w = A z the Krylov iterates of standard (preconditioned) CG follow from ONE fused reduction of (r.z, w.z, r.r):
//     beta = gamma / gamma_old,   alpha = gamma / (delta - beta gamma / alpha_old)
//     p = z + beta p,  s = w + beta s,  x += alpha p,  r -= alpha s,  z = M^-1 r         (s tracks A p)
// r, s (and z with Jacobi) are double buffered so that the halo push can recompute a neighbour's z_j from values no
// thread overwrites.  Workspace W = [r0 r1 | s0 s1 | z0 z1 | p | w].
template <bool PEER, bool C16>
__global__ void __launch_bounds__(PT, 1)
k_cg1_sell(KrylovArgs K, SellDev S, RedDev R, PeerView pv) {
    __shared__ double sh[NR * 32];
    __shared__ double s_red[9][NR];
    __shared__ int s_last;
    const int n = K.n, ld = K.ld;
    const int gtid = blockIdx.x * PT + threadIdx.x, gsz = gridDim.x * PT;
    double *rbuf = K.W, *sbuf = K.W + 2 * (size_t)ld, *zbuf = K.W + 4 * (size_t)ld, *p = K.W + 6 * (size_t)ld,
           *w = K.W + 7 * (size_t)ld;
    const double* dinv = K.dinv;
    const double* b = K.b;
    double* x = K.x;
    unsigned tag = pv.tag0;
    unsigned halo_tag = 0;
    int halo_buf = 0;
    const LLWord* halo0 = PEER ? pv.my_halo : nullptr;

    // ---- r0 = b - A x0, z0 = M^-1 r0, p = s = 0 ---------------------------------------------------------------------------
    if (PEER) {
        ++tag;
        push_halo(pv, 0, tag, gtid, gsz, [&](int j) { return x[j]; });
        halo_tag = tag; halo_buf = 0;
    }
    unsigned nbar = 0;
    spmv_sell<PEER, C16>(S, n, x, w, halo0, halo_tag, pv.error, [](int, double) {});
    grid_barrier_local(R, nbar);   // w complete before it is read
    int cur = 0;
    double* r = rbuf;
    const double* z = dinv ? zbuf : rbuf;
    double gam_p = 0, rr_p = 0, bb = 0;            // per-thread partials of r.z, r.r (carried to the next reduction)
    for (int i = gtid; i < n; i += gsz) {
        const double bi = b[i], ri = bi - w[i];
        const double zi = dinv ? dinv[i] * ri : ri;
        r[i] = ri;
        if (dinv) zbuf[i] = zi;
        p[i] = 0.0;
        sbuf[i] = 0.0;
        gam_p += ri * zi; rr_p += ri * ri; bb += bi * bi;
    }
    if (PEER) {
        ++tag;
        push_halo(pv, 1, tag, gtid, gsz, [&](int j) { const double ri = b[j] - w[j]; return dinv ? dinv[j] * ri : ri; });
        halo_tag = tag; halo_buf = 1;
    }
    double rr;
    {
        double v[NR] = {gam_p, rr_p, bb, 0};
        ++tag;
        grid_sum4<PEER>(v, R, tag, pv, sh, s_red, &s_last);
        rr = v[1]; bb = v[2];
    }
    const double thr = K.rtol * K.rtol * bb;
    int it = 0;
    bool conv = (rr <= thr) || (bb == 0.0), bad = !isfinite(rr) || !isfinite(bb);
    double gam_old = 1.0, alpha = 1.0;

    while (!conv && !bad && it < K.maxit) {
        // ---- w = A z, then ONE reduction of (r.z, w.z, r.r) -----------------------------------------------------------------
        double del = 0;
        FDB_STAMP(0);
        spmv_sell<PEER, C16>(S, n, z, w, halo0 + (size_t)halo_buf * pv.n_halo, halo_tag, pv.error,
                             [&](int row, double s) { del += s * z[row]; });
        FDB_STAMP(1);
        double v[NR] = {gam_p, del, rr_p, 0};
        ++tag;
        const int err = grid_sum4<PEER>(v, R, tag, pv, sh, s_red, &s_last);
        FDB_STAMP(2);
        const double gam = v[0];
        del = v[1]; rr = v[2];
        if (!isfinite(rr) || !isfinite(del) || err) { bad = true; break; }   // same sums everywhere: a uniform exit
        if (it > 0 && gtid == 0) K.hist[(it - 1) % K.hist_cap] = rr;
        conv = rr <= thr;
        if (conv) break;
        const double beta = it == 0 ? 0.0 : gam / gam_old;
        alpha = it == 0 ? gam / del : gam / (del - beta * gam / alpha);
        gam_old = gam;
        // ---- p, s, x, r, z; partials of the next reduction; halo push of the new z -----------------------------------------
        const double* r_old = rbuf + (size_t)cur * ld;
        const double* s_old = sbuf + (size_t)cur * ld;
        double* r_new = rbuf + (size_t)(cur ^ 1) * ld;
        double* s_new = sbuf + (size_t)(cur ^ 1) * ld;
        double* z_new = dinv ? zbuf + (size_t)(cur ^ 1) * ld : r_new;
        gam_p = 0; rr_p = 0;
        for (int i = gtid; i < n; i += gsz) {
            const double zi = z[i];
            const double pi = zi + beta * p[i];
            const double si = w[i] + beta * s_old[i];
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
            push_halo(pv, halo_buf ^ 1, tag, gtid, gsz, [&](int j) {
                const double ri = r_old[j] - alpha * (w[j] + beta * s_old[j]);
                return dinv ? dinv[j] * ri : ri;
            });
            halo_tag = tag; halo_buf ^= 1;
        }
        cur ^= 1;
        z = z_new;
        FDB_STAMP(3);
        grid_barrier_local(R, nbar);   // z complete before the next SpMV
        FDB_STAMP(4);
        ++it;
    }
    if (!conv && !bad && it >= K.maxit) {   // budget exhausted: the residual of the last update has not been reduced yet
        double v[NR] = {gam_p, 0, rr_p, 0};
        ++tag;
        grid_sum4<PEER>(v, R, tag, pv, sh, s_red, &s_last);
        rr = v[2];
        conv = rr <= thr;
    }
    if (gtid == 0) {
        K.sc->pad = (int)tag;   // last tag used: the host carries it to the next solve on this matrix
        K.sc->bb = bb; K.sc->thr = thr; K.sc->rr = rr; K.sc->iters = it;
        K.sc->done = conv ? 1 : 0;
        K.sc->breakdown = (bb == 0.0) ? 2 : (bad ? 3 : 0);
    }
}

// ---- BiCGSTAB ---------------------------------------------------------------------------------------------------------------
// Same recurrence, breakdown tests (Eigen's: |rho| < eps^2 |r0|^2, r0.v = 0, omega = 0) and restart rule as the multi-kernel
// loop of solve.cu: on a breakdown nothing is updated and the iteration restarts from the current x with a fresh shadow
// residual.  Workspace W = [r | r0 | p0 | p1 | v | s | t | y | z]; y = M^-1 p and z = M^-1 s alias p and s without
// Jacobi; p is double buffered so that the halo push can recompute a neighbour's y_j from values no thread overwrites.
// Per iteration: (p, y) -> barrier -> v = A y, r0.v -> reduce -> (s, z) -> barrier -> t = A z, t.t, t.s -> reduce ->
// (x, r), r.r, r0.r -> reduce.
template <bool PEER, bool C16>
__global__ void __launch_bounds__(PT, 1)
k_bicgstab_sell(KrylovArgs K, SellDev S, RedDev R, PeerView pv) {
    __shared__ double sh[NR * 32];
    __shared__ double s_red[9][NR];
    __shared__ int s_last;
    const int n = K.n;
    const size_t ld = (size_t)K.ld;
    const int gtid = blockIdx.x * PT + threadIdx.x, gsz = gridDim.x * PT;
    double *r = K.W, *r0 = K.W + ld, *pbuf = K.W + 2 * ld, *vv = K.W + 4 * ld, *sv = K.W + 5 * ld, *tv = K.W + 6 * ld;
    const double* dinv = K.dinv;
    double* ybuf = K.W + 7 * ld;
    double* zs = dinv ? K.W + 8 * ld : sv;
    int pc = 0;                                  // which p buffer holds the current direction
    const double* b = K.b;
    double* x = K.x;
    unsigned tag = pv.tag0;
    const LLWord* halo0 = PEER ? pv.my_halo : nullptr;
    int hb = 0;                                  // halo receive buffer of the next push (alternates)
    const double eps = 2.220446049250313e-16;
    unsigned nbar = 0;
    auto barrier = [&]() { grid_barrier_local(R, nbar); };
    int it = 0, restarts = 0;
    double bb = 0, thr = 0, rr = 0;
    bool conv = false, bad = false;
    for (;;) {   // (re)start from the current x
        unsigned ht = 0;
        if (PEER) {
            ++tag;
            push_halo(pv, hb, tag, gtid, gsz, [&](int j) { return x[j]; });
            ht = tag;
        }
        spmv_sell<PEER, C16>(S, n, x, vv, halo0 + (size_t)hb * pv.n_halo, ht, pv.error, [](int, double) {});
        hb ^= 1;
        barrier();
        double rr_p = 0, bb_p = 0;
        for (int i = gtid; i < n; i += gsz) {
            const double bi = b[i], ri = bi - vv[i];
            r[i] = ri; r0[i] = ri;
            rr_p += ri * ri; bb_p += bi * bi;
        }
        double rho, r0sq;
        {
            double v[NR] = {rr_p, bb_p, 0, 0};
            ++tag;
            grid_sum4<PEER>(v, R, tag, pv, sh, s_red, &s_last);
            rr = v[0];
            if (restarts == 0) { bb = v[1]; thr = K.rtol * K.rtol * bb; }
        }
        rho = rr; r0sq = rr;
        conv = (rr <= thr) || (bb == 0.0);
        bad = !isfinite(rr) || !isfinite(bb);
        double rho_old = 1, alpha = 1, omega = 1;
        bool broke = false, first = true;
        while (!conv && !bad && it < K.maxit) {
            // ---- beta, p, y = M^-1 p (+ halo push of y) -----------------------------------------------------------------------
            if (!(fabs(rho) >= eps * eps * r0sq) || !isfinite(rho)) { broke = true; break; }
            const double beta = first ? 0.0 : (rho / rho_old) * (alpha / omega);
            const double* p_old = pbuf + (size_t)pc * ld;
            double* p_new = pbuf + (size_t)(pc ^ 1) * ld;
            double* y = dinv ? ybuf : p_new;
            for (int i = gtid; i < n; i += gsz) {
                const double pi = first ? r[i] : r[i] + beta * (p_old[i] - omega * vv[i]);
                p_new[i] = pi;
                if (dinv) y[i] = dinv[i] * pi;
            }
            if (PEER) {
                ++tag;
                push_halo(pv, hb, tag, gtid, gsz, [&](int j) {
                    const double pj = first ? r[j] : r[j] + beta * (p_old[j] - omega * vv[j]);
                    return dinv ? dinv[j] * pj : pj;
                });
                ht = tag;
            }
            pc ^= 1;
            barrier();
            // ---- v = A y, r0.v ---------------------------------------------------------------------------------------------------
            double r0v = 0;
            spmv_sell<PEER, C16>(S, n, y, vv, halo0 + (size_t)hb * pv.n_halo, ht, pv.error,
                                 [&](int row, double s) { r0v += s * r0[row]; });
            hb ^= 1;
            {
                double v[NR] = {r0v, 0, 0, 0};
                ++tag;
                grid_sum4<PEER>(v, R, tag, pv, sh, s_red, &s_last);
                r0v = v[0];
            }
            alpha = rho / r0v;
            if (!isfinite(alpha)) { broke = true; break; }
            // ---- s = r - alpha v, z = M^-1 s (+ halo push of z) --------------------------------------------------------------------
            for (int i = gtid; i < n; i += gsz) {
                const double si = r[i] - alpha * vv[i];
                sv[i] = si;
                if (dinv) zs[i] = dinv[i] * si;
            }
            if (PEER) {
                ++tag;
                push_halo(pv, hb, tag, gtid, gsz, [&](int j) {
                    const double sj = r[j] - alpha * vv[j];
                    return dinv ? dinv[j] * sj : sj;
                });
                ht = tag;
            }
            barrier();
            // ---- t = A z, t.t, t.s ------------------------------------------------------------------------------------------------
            double tt = 0, ts = 0;
            spmv_sell<PEER, C16>(S, n, zs, tv, halo0 + (size_t)hb * pv.n_halo, ht, pv.error,
                                 [&](int row, double s) { tt += s * s; ts += s * sv[row]; });
            hb ^= 1;
            {
                double v[NR] = {tt, ts, 0, 0};
                ++tag;
                grid_sum4<PEER>(v, R, tag, pv, sh, s_red, &s_last);
                tt = v[0]; ts = v[1];
            }
            omega = tt > 0 ? ts / tt : 0.0;
            if (!isfinite(omega)) { broke = true; break; }
            // ---- x += alpha y + omega z, r = s - omega t, r.r, r0.r ----------------------------------------------------------------
            double rr_n = 0, rho_n = 0;
            for (int i = gtid; i < n; i += gsz) {
                x[i] += alpha * y[i] + omega * zs[i];
                const double ri = sv[i] - omega * tv[i];
                r[i] = ri;
                rr_n += ri * ri;
                rho_n += r0[i] * ri;
            }
            {
                double v[NR] = {rr_n, rho_n, 0, 0};
                ++tag;
                const int err = grid_sum4<PEER>(v, R, tag, pv, sh, s_red, &s_last);
                rr = v[0]; rho_n = v[1];
                if (err) bad = true;
            }
            if (gtid == 0) K.hist[it % K.hist_cap] = rr;
            ++it;
            first = false;
            rho_old = rho;
            rho = rho_n;
            conv = rr <= thr;
            if (!isfinite(rr) || bad) { bad = true; break; }
            if (omega == 0.0 && !conv) { broke = true; break; }
        }
        if (!broke || conv || bad || restarts >= 16 || it >= K.maxit) break;
        ++restarts;
    }
    if (gtid == 0) {
        K.sc->pad = (int)tag;
        K.sc->bb = bb; K.sc->thr = thr; K.sc->rr = rr; K.sc->iters = it;
        K.sc->done = conv ? 1 : 0;
        K.sc->breakdown = (bb == 0.0) ? 2 : (bad ? 3 : 0);
    }
}

// ===========================================================================================================================
__global__ void k_slice_halo_flags(int n_owned, int n_slices, const int32_t* __restrict__ rowptr,
                                   const int32_t* __restrict__ colidx, uint8_t* __restrict__ flag) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= n_slices) return;
    const int row = 32 * s + lane;
    int touches = 0;
    if (row < n_owned)
        for (int t = rowptr[row]; t < rowptr[row + 1]; ++t) touches |= (colidx[t] >= n_owned);
    touches = __any_sync(0xffffffffu, touches);
    if (lane == 0) flag[s] = (uint8_t)touches;
}

__global__ void k_jacobi_peer(int n, const int32_t* __restrict__ diag, const double* __restrict__ val,
                              double* __restrict__ dinv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = diag[i];
    double a = d >= 0 ? val[d] : 0.0;
    dinv[i] = a != 0.0 ? 1.0 / a : 1.0;
}

// visiting order of the owned slices of a partitioned matrix: interior slices first, halo-coupled slices last
static int ensure_slice_order(fdb_matrix* A, int n_owned) {
    if (A->slice_order.n > 0 && A->slice_order_n == n_owned) return FDB_OK;
    fdb_space* s = A->space;
    const Pattern* P = A->pat;
    const int n_slices = (n_owned + 31) / 32;
    DevBuf<uint8_t> flag;
    FDB_TRY(flag.alloc(n_slices));
    k_slice_halo_flags<<<(n_slices + 7) / 8, 256, 0, s->stream>>>(n_owned, n_slices, P->rowptr.p, P->colidx.p, flag.p);
    FDB_CUDA(cudaGetLastError());
    std::vector<uint8_t> h(n_slices);
    FDB_CUDA(cudaMemcpyAsync(h.data(), flag.p, n_slices, cudaMemcpyDeviceToHost, s->stream));
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    std::vector<int32_t> order;
    order.reserve(n_slices);
    for (int k = 0; k < n_slices; ++k) if (!h[k]) order.push_back(k);
    for (int k = 0; k < n_slices; ++k) if (h[k]) order.push_back(k);
    FDB_TRY(A->slice_order.alloc(n_slices));
    FDB_CUDA(cudaMemcpyAsync(A->slice_order.p, order.data(), sizeof(int32_t) * n_slices, cudaMemcpyHostToDevice, s->stream));
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    A->slice_order_n = n_owned;
    return FDB_OK;
}

int& peer_sell_mode() {
    static int mode = -1;   // -1: from the environment; 0: off; 1: partitioned matrices with a peer plan (default); 2: always
    return mode;
}

// Runs CG / BiCGSTAB as one persistent sliced-ELL kernel when the matrix admits it; *handled tells the caller.
int solve_persistent_sell(fdb_matrix* A, const double* b, double* x, const fdb_solver_opts* o, fdb_solve_stats* stats,
                          bool* handled) {
    *handled = false;
    fdb_space* s = A->space;
    Partition* part = A->part;
    int& mode = peer_sell_mode();
    if (mode < 0) mode = getenv("FDB_NO_PEER_SELL") ? 0 : (getenv("FDB_PEER_SELL_ALWAYS") ? 2 : 1);
    if (mode == 0) return FDB_OK;
    if (!part && mode != 2) return FDB_OK;
    if (part && !part->peer_ready) return FDB_OK;   // partitioned without a peer-memory plan: NCCL multi-kernel loop
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s->device);
    if (!coop) return FDB_OK;
    Pattern* P = const_cast<Pattern*>(A->pat);
    if (!sell_view_ready(A)) return FDB_OK;         // padding too large: keep the CSR kernels
    cudaStream_t st = s->stream;
    const bool peer = part != nullptr;
    const bool cg = o->kind == FDB_SOLVER_CG;
    const size_t ld = (size_t)s->n_dofs;
    const int n = part ? part->n_owned : s->n_dofs;
    const long long n_glob = part ? part->n_global : (long long)s->n_dofs;   // identical on every rank
    const int maxit = o->maxit > 0 ? o->maxit : (int)std::min<long long>(10 * std::max<long long>(n_glob, 1), 2000000000LL);
    const bool jac = o->jacobi != 0;
    const int hist_cap = 1 << 16;
    const bool c16 = P->col16_state > 0;
    const int grid = s->sm_count;                   // one block per SM (cooperative launch: all co-resident)
    const size_t nvec = 9;   // CG1 uses 8 of them, BiCGSTAB 9; + the Jacobi diagonal
    if (A->work.n < (nvec + 1) * ld) FDB_TRY(A->work.alloc((nvec + 1) * ld));
    const size_t part_doubles = (size_t)NR * grid + 64;
    if (A->partials.n < part_doubles + 64) FDB_TRY(A->partials.alloc(part_doubles + 64));
    if (A->hist.n < (size_t)hist_cap) FDB_TRY(A->hist.alloc((size_t)hist_cap));
    double* PA = A->partials.p;
    Scal* sc = reinterpret_cast<Scal*>(PA + (size_t)NR * grid);
    unsigned* ticket = reinterpret_cast<unsigned*>(PA + (size_t)NR * grid + 16);
    FDB_CUDA(cudaMemsetAsync(ticket, 0, 64, st));
    double* dinv = A->work.p + nvec * ld;
    if (jac) {
        k_jacobi_peer<<<(n + 255) / 256, 256, 0, st>>>(n, P->diag.p, A->val.p, dinv);
        FDB_CUDA(cudaGetLastError());
    }
    if (peer) FDB_TRY(ensure_slice_order(A, n));
    PeerView pv;
    memset(&pv, 0, sizeof(pv));
    if (peer) {
        pv = *reinterpret_cast<PeerView*>(part->peer_view);
        pv.tag0 = part->peer_tag;   // every rank runs the same sequence of solves on this matrix
        FDB_CUDA(cudaMemsetAsync(pv.error, 0, sizeof(int), st));
    } else {   // one rank: the reduction lines live in this matrix's own scratch
        if (A->local_red.n < 4 * NR + 8) {
            FDB_TRY(A->local_red.alloc(4 * NR + 8));
            FDB_CUDA(cudaMemsetAsync(A->local_red.p, 0, sizeof(LLWord) * (4 * NR + 8), st));
            A->local_tag = 0;
        }
        pv.world = 1; pv.rank = 0;
        pv.red_of[0] = pv.my_red = A->local_red.p;
        pv.error = reinterpret_cast<int*>(A->local_red.p + 4 * NR);
        pv.tag0 = A->local_tag;
        FDB_CUDA(cudaMemsetAsync(pv.error, 0, sizeof(int), st));
    }
    static DevBuf<unsigned long long> trace_buf;
    const bool tracing = getenv("FDB_CG_TRACE") != nullptr && cg;
    if (tracing) {
        if (!trace_buf.p) FDB_TRY(trace_buf.alloc(64 * 8));
        FDB_CUDA(cudaMemsetAsync(trace_buf.p, 0, 64 * 8 * 8, st));
    }
    KrylovArgs K{n, (int)ld, b, x, A->work.p, jac ? dinv : nullptr, sc, A->hist.p, hist_cap, maxit, o->rtol,
                 tracing ? trace_buf.p : nullptr};
    SellDev S{P->sell_ptr.p, c16 ? static_cast<const void*>(P->sell_col16.p) : static_cast<const void*>(P->sell_col.p),
              A->sell_val.p, peer ? A->slice_order.p : nullptr, (n + 31) / 32};
    RedDev R{PA, ticket, ticket + 1, ticket + 2};
    cudaEvent_t ev0, ev1;
    FDB_CUDA(cudaEventCreate(&ev0));
    FDB_CUDA(cudaEventCreate(&ev1));
    FDB_CUDA(cudaEventRecord(ev0, st));
    void* args[] = {&K, &S, &R, &pv};
    const void* fn;
    if (cg) fn = peer ? (c16 ? (const void*)k_cg1_sell<true, true> : (const void*)k_cg1_sell<true, false>)
                      : (c16 ? (const void*)k_cg1_sell<false, true> : (const void*)k_cg1_sell<false, false>);
    else fn = peer ? (c16 ? (const void*)k_bicgstab_sell<true, true> : (const void*)k_bicgstab_sell<true, false>)
                   : (c16 ? (const void*)k_bicgstab_sell<false, true> : (const void*)k_bicgstab_sell<false, false>);
    FDB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(PT), args, 0, st));
    FDB_CUDA(cudaEventRecord(ev1, st));
    Scal h;
    FDB_CUDA(cudaMemcpyAsync(&h, sc, sizeof(Scal), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (peer) part->peer_tag = (unsigned)h.pad;
    else A->local_tag = (unsigned)h.pad;
    if (tracing) {   // average phase durations over iterations 8..62 as block 0 saw them
        std::vector<unsigned long long> t(64 * 8);
        FDB_CUDA(cudaMemcpy(t.data(), trace_buf.p, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost));
        double acc[5] = {0, 0, 0, 0, 0};
        int cnt = 0;
        for (int it = 8; it < 62; ++it) {
            if (!t[it * 8 + 4] || !t[(it + 1) * 8]) break;
            for (int k = 0; k < 4; ++k) acc[k] += (double)(t[it * 8 + k + 1] - t[it * 8 + k]);
            acc[4] += (double)(t[(it + 1) * 8] - t[it * 8]);
            ++cnt;
        }
        if (cnt)
            fprintf(stderr, "[fdb] CG1 trace (us, %d iterations, block 0): spmv %.1f | reduce %.1f | update+push %.1f | barrier %.1f | "
                    "iteration %.1f\n", cnt, acc[0] / cnt / 1e3, acc[1] / cnt / 1e3, acc[2] / cnt / 1e3, acc[3] / cnt / 1e3,
                    acc[4] / cnt / 1e3);
    }
    {
        int err = 0;
        FDB_CUDA(cudaMemcpy(&err, pv.error, sizeof(int), cudaMemcpyDeviceToHost));
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
