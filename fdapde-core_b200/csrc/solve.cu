// K7/K8: CSR SpMV and the fused Krylov solvers that replace the reference's direct solve
//   FEMLinearEllipticSolver::solve   finite_elements/solvers/fem_linear_elliptic_solver.h:34-50
//   (Eigen::SparseLU<SpMatrix<double>, COLAMDOrdering<int>>::compute / solve)
// The reference has no iterative solver; CG (SPD operators) and BiCGSTAB (non-symmetric) are
// judged against the direct solution.  Design:
//   * SpMV: CSR, a power-of-two group of threads per row chosen from nnz/row (vector-row .. warp-per-row),
//     warp-shuffle reduction inside the group, persistent grid-stride blocks.
//   * dot products are fused into the kernel that produces their operands: warp shuffle -> block -> one partial
//     per block; every consumer block re-sums the (few hundred) partials in a fixed order, so there are no
//     atomics, no extra reduction launches and results are bit-reproducible.
//   * scalars (alpha, beta, omega, rho) never visit the host; convergence is detected on the device, iterations
//     launched after it are no-ops, the host only polls a flag every `check_every` iterations.
#include <cub/cub.cuh>

#include <algorithm>

#include "solve_common.cuh"

namespace fdb {

// ---- K7: y = A x (+ optional fused dot w.y) -----------------------------------------------------------------------
// column of stored entry t of `row`: 32-bit index, or -- when every |col - row| of the pattern fits -- a 16-bit offset
// from the row (10 instead of 12 bytes per stored entry: SpMV is bound by exactly this stream)
template <bool C16>
__device__ __forceinline__ int column_of(const void* __restrict__ cols, int t, int row) {
    if constexpr (C16) return row + (int)__ldg(static_cast<const int16_t*>(cols) + t);
    else return __ldg(static_cast<const int32_t*>(cols) + t);
}

template <int TPR, bool DOT, bool C16>
__global__ void __launch_bounds__(VB)
k_spmv(int n, const int32_t* __restrict__ rowptr, const void* __restrict__ colidx, const double* __restrict__ val,
       const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ w, double* __restrict__ part,
       const int* __restrict__ done) {
    __shared__ double sh[VB / 32];
    if (done && *done) return;
    constexpr int RPB = VB / TPR;
    const int lane = threadIdx.x % TPR, rl = threadIdx.x / TPR;
    double local = 0;
    for (int base = blockIdx.x * RPB; base < n; base += gridDim.x * RPB) {
        int row = base + rl;
        double s = 0;
        if (row < n) {
            int t0 = rowptr[row], t1 = rowptr[row + 1];
            for (int t = t0 + lane; t < t1; t += TPR) s += val[t] * __ldg(x + column_of<C16>(colidx, t, row));
        }
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (row < n && lane == 0) {
            y[row] = s;
            if (DOT) local += s * w[row];
        }
    }
    if (DOT) {
        double t = block_sum(local, sh);
        if (threadIdx.x == 0) part[blockIdx.x] = t;
    }
}

// ---- CG -------------------------------------------------------------------------------------------------------------
// r = b - q ; z = dinv r ; p = z ; partials: rz, rr, bb
__global__ void __launch_bounds__(VB)
k_cg_init(int n, const double* __restrict__ b, const double* __restrict__ q, const double* __restrict__ dinv,
          double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, double* __restrict__ part_rz,
          double* __restrict__ part_rr, double* __restrict__ part_bb) {
    __shared__ double sh[VB / 32];
    double rz = 0, rr = 0, bb = 0;
    for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
        double bi = b[i], ri = bi - q[i];
        double zi = dinv ? dinv[i] * ri : ri;
        r[i] = ri;
        if (dinv) z[i] = zi;
        p[i] = zi;
        rz += ri * zi;
        rr += ri * ri;
        bb += bi * bi;
    }
    rz = block_sum(rz, sh);
    rr = block_sum(rr, sh);
    bb = block_sum(bb, sh);
    if (threadIdx.x == 0) { part_rz[blockIdx.x] = rz; part_rr[blockIdx.x] = rr; part_bb[blockIdx.x] = bb; }
}

__global__ void __launch_bounds__(VB)
k_set_threshold(int np, const double* __restrict__ part_rr, const double* __restrict__ part_bb, double rtol,
                Scal* sc, int restart = 0) {
    __shared__ double sh[VB / 32];
    double rr = sum_partials(part_rr, np, sh);
    double bb = sum_partials(part_bb, np, sh);
    if (threadIdx.x == 0) {
        sc->bb = bb;
        sc->thr = rtol * rtol * bb;
        sc->rr = rr;
        sc->r0sq = rr;
        sc->done = (rr <= sc->thr || bb == 0.0) ? 1 : 0;
        if (!restart) sc->iters = 0;  // a BiCGSTAB restart keeps counting
        sc->breakdown = (bb == 0.0) ? 2 : 0;  // 2: zero right-hand side => x = 0 (Eigen's convention)
        sc->rho = 1; sc->alpha = 1; sc->omega = 1;
    }
}

// alpha = rz / (p.q) ; x += alpha p ; r -= alpha q ; z = dinv r ; partials rz_new, rr_new
__global__ void __launch_bounds__(VB)
k_cg_update(int n, int np, const double* __restrict__ part_pq, const double* __restrict__ part_rz_old,
            const double* __restrict__ p, const double* __restrict__ q, const double* __restrict__ dinv,
            double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, double* __restrict__ part_rz_new,
            double* __restrict__ part_rr_new, const Scal* __restrict__ sc) {
    __shared__ double sh[3 * (VB / 32)];
    if (sc->done) return;
    double pq, rz_old, unused;
    sum_partials3(part_pq, part_rz_old, nullptr, np, sh, pq, rz_old, unused);
    double alpha = rz_old / pq;
    double rz = 0, rr = 0;
    for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
        x[i] += alpha * p[i];
        double ri = r[i] - alpha * q[i];
        r[i] = ri;
        double zi = dinv ? dinv[i] * ri : ri;
        if (dinv) z[i] = zi;
        rz += ri * zi;
        rr += ri * ri;
    }
    rz = block_sum(rz, sh);
    rr = block_sum(rr, sh);
    if (threadIdx.x == 0) { part_rz_new[blockIdx.x] = rz; part_rr_new[blockIdx.x] = rr; }
}

// beta = rz_new / rz_old ; p = z + beta p ; records the residual and raises `done`.
// Only block 0 / thread 0 reads or writes sc->done here; every block decides convergence from the same partials,
// so a block can never observe `done` flipping inside this kernel.  After convergence k_spmv / k_cg_update return
// early, the partials stay frozen, and this kernel keeps evaluating conv == true (no update, no bookkeeping).
__global__ void __launch_bounds__(VB)
k_cg_direction(int n, int np, const double* __restrict__ part_rz_new, const double* __restrict__ part_rz_old,
               const double* __restrict__ part_rr_new, const double* __restrict__ z, double* __restrict__ p, Scal* sc,
               double* __restrict__ hist, int maxit, int hist_cap) {
    __shared__ double sh[3 * (VB / 32)];
    double rr, rz_new, rz_old;
    sum_partials3(part_rr_new, part_rz_new, part_rz_old, np, sh, rr, rz_new, rz_old);
    const bool conv = rr <= sc->thr || !isfinite(rr);   // a non-finite residual ends the solve (reported as not converged)
    if (!conv) {
        double beta = rz_new / rz_old;
        for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) p[i] = z[i] + beta * p[i];
    }
    // the iteration number lives on the device (sc->iters), so the same launch sequence can be replayed from a
    // CUDA graph; `done` also stops the replay at the iteration budget
    if (blockIdx.x == 0 && threadIdx.x == 0 && !sc->done) {
        const int it = sc->iters;
        hist[it % hist_cap] = rr;
        sc->rr = rr;
        sc->iters = it + 1;
        if (conv || it + 1 >= maxit) sc->done = 1;
        if (!isfinite(rr)) sc->breakdown = 3;
    }
}

// ---- BiCGSTAB -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VB)
k_bi_init(int n, const double* __restrict__ b, const double* __restrict__ q, double* __restrict__ r,
          double* __restrict__ r0, double* __restrict__ p, double* __restrict__ v, double* __restrict__ part_rho,
          double* __restrict__ part_rr, double* __restrict__ part_bb) {
    __shared__ double sh[VB / 32];
    double rr = 0, bb = 0;
    for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
        double bi = b[i], ri = bi - q[i];
        r[i] = ri; r0[i] = ri; p[i] = 0; v[i] = 0;
        rr += ri * ri;
        bb += bi * bi;
    }
    rr = block_sum(rr, sh);
    bb = block_sum(bb, sh);
    if (threadIdx.x == 0) { part_rho[blockIdx.x] = rr; part_rr[blockIdx.x] = rr; part_bb[blockIdx.x] = bb; }
}

// rho_new = r0.r (partials) ; beta = (rho_new/rho)(alpha/omega) ; p = r + beta (p - omega v) ; y = dinv p
__global__ void __launch_bounds__(VB)
k_bi_p(int n, int np, int first, const double* __restrict__ part_rho, const double* __restrict__ r,
       const double* __restrict__ v, const double* __restrict__ dinv, double* __restrict__ p, double* __restrict__ y,
       Scal* sc) {
    __shared__ double sh[VB / 32];
    if (sc->done) return;
    double rho_new = sum_partials(part_rho, np, sh);
    double rho = sc->rho, alpha = sc->alpha, omega = sc->omega;
    __syncthreads();
    // breakdown test of Eigen's BiCGSTAB (|rho| < eps^2 |r0|^2): the shadow residual has become orthogonal to r.
    // Nothing is updated; the host restarts from the current x with a fresh shadow residual.
    const double eps = 2.220446049250313e-16;
    const bool bad = !(fabs(rho_new) >= eps * eps * sc->r0sq) || !isfinite(rho_new);
    __syncthreads();
    if (bad) {
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->breakdown = 1;
        return;
    }
    double beta = first ? 0.0 : (rho_new / rho) * (alpha / omega);
    for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
        double pi = first ? r[i] : r[i] + beta * (p[i] - omega * v[i]);
        p[i] = pi;
        if (dinv) y[i] = dinv[i] * pi;
    }
}
__global__ void k_bi_store_rho(int np, const double* __restrict__ part_rho, Scal* sc) {
    __shared__ double sh[VB / 32];
    if (sc->done) return;
    double rho_new = sum_partials(part_rho, np, sh);
    if (threadIdx.x == 0) {
        sc->rho = rho_new;
        if (sc->breakdown == 1) sc->done = 1;  // raised by k_bi_p: every later kernel of this iteration is a no-op
    }
}

// alpha = rho / (r0.v) ; s = r - alpha v ; z = dinv s
__global__ void __launch_bounds__(VB)
k_bi_s(int n, int np, const double* __restrict__ part_r0v, const double* __restrict__ r, const double* __restrict__ v,
       const double* __restrict__ dinv, double* __restrict__ s, double* __restrict__ z, Scal* sc) {
    __shared__ double sh[VB / 32];
    if (sc->done) return;
    double r0v = sum_partials(part_r0v, np, sh);
    double alpha = sc->rho / r0v;
    if (!isfinite(alpha)) {  // r0.v == 0: breakdown, leave x untouched (the host restarts)
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->breakdown = 1;
        return;
    }
    for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
        double si = r[i] - alpha * v[i];
        s[i] = si;
        if (dinv) z[i] = dinv[i] * si;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) sc->alpha = alpha;
}

// t = A z fused with partials t.t and t.s  (two dots => dedicated SpMV variant)
template <int TPR, bool C16>
__global__ void __launch_bounds__(VB)
k_spmv_tt_ts(int n, const int32_t* __restrict__ rowptr, const void* __restrict__ colidx,
             const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
             const double* __restrict__ s, double* __restrict__ part_tt, double* __restrict__ part_ts,
             const Scal* __restrict__ sc) {
    __shared__ double sh[VB / 32];
    if (sc->done) return;
    constexpr int RPB = VB / TPR;
    const int lane = threadIdx.x % TPR, rl = threadIdx.x / TPR;
    double tt = 0, ts = 0;
    for (int base = blockIdx.x * RPB; base < n; base += gridDim.x * RPB) {
        int row = base + rl;
        double a = 0;
        if (row < n) {
            int t0 = rowptr[row], t1 = rowptr[row + 1];
            for (int t = t0 + lane; t < t1; t += TPR) a += val[t] * __ldg(x + column_of<C16>(colidx, t, row));
        }
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (row < n && lane == 0) {
            y[row] = a;
            tt += a * a;
            ts += a * s[row];
        }
    }
    tt = block_sum(tt, sh);
    ts = block_sum(ts, sh);
    if (threadIdx.x == 0) { part_tt[blockIdx.x] = tt; part_ts[blockIdx.x] = ts; }
}

// omega = (t.s)/(t.t) ; x += alpha y + omega z ; r = s - omega t ; partials rr, r0.r
__global__ void __launch_bounds__(VB)
k_bi_x(int n, int np, const double* __restrict__ part_tt, const double* __restrict__ part_ts,
       const double* __restrict__ y, const double* __restrict__ z, const double* __restrict__ s,
       const double* __restrict__ t, const double* __restrict__ r0, double* __restrict__ x, double* __restrict__ r,
       double* __restrict__ part_rr, double* __restrict__ part_rho, Scal* sc) {
    __shared__ double sh[VB / 32];
    if (sc->done) return;
    double tt = sum_partials(part_tt, np, sh);
    double ts = sum_partials(part_ts, np, sh);
    double omega = tt > 0 ? ts / tt : 0.0;
    double alpha = sc->alpha;
    const int broke = sc->breakdown;
    __syncthreads();
    if (broke == 1 || !isfinite(omega) || !isfinite(alpha)) {  // earlier breakdown in this iteration: no update
        if (blockIdx.x == 0 && threadIdx.x == 0) sc->breakdown = 1;
        return;
    }
    double rr = 0, rho = 0;
    for (int i = blockIdx.x * VB + threadIdx.x; i < n; i += gridDim.x * VB) {
        x[i] += alpha * y[i] + omega * z[i];
        double ri = s[i] - omega * t[i];
        r[i] = ri;
        rr += ri * ri;
        rho += r0[i] * ri;
    }
    rr = block_sum(rr, sh);
    rho = block_sum(rho, sh);
    if (threadIdx.x == 0) { part_rr[blockIdx.x] = rr; part_rho[blockIdx.x] = rho; }
    // sc->omega and the omega == 0 breakdown flag are written by k_bi_finish (single block, next launch): written here
    // by block 0 they could be seen by a late-scheduled block of this same grid, which would then skip its slice
}
__global__ void k_bi_finish(int np, const double* __restrict__ part_rr, const double* __restrict__ part_tt,
                            const double* __restrict__ part_ts, Scal* sc, double* __restrict__ hist, int it,
                            int hist_cap) {
    __shared__ double sh[VB / 32];
    if (sc->done) return;
    double rr = sum_partials(part_rr, np, sh);
    const double tt = sum_partials(part_tt, np, sh);
    const double ts = sum_partials(part_ts, np, sh);
    if (threadIdx.x == 0) {
        if (sc->breakdown != 1) {   // same omega as every block of k_bi_x computed
            const double omega = tt > 0 ? ts / tt : 0.0;
            sc->omega = omega;
            if (omega == 0.0) sc->breakdown = 1;
        }
        hist[it % hist_cap] = rr;
        sc->rr = rr;
        sc->iters = it + 1;
        if (rr <= sc->thr || sc->breakdown) sc->done = 1;
    }
}

__global__ void k_jacobi(int n, const int32_t* __restrict__ diag, const double* __restrict__ val,
                         double* __restrict__ dinv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = diag[i];
    double a = d >= 0 ? val[d] : 0.0;
    dinv[i] = a != 0.0 ? 1.0 / a : 1.0;
}

// =====================================================================================================================
int pick_tpr(const fdb::Pattern* P, int n) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("FDB_SPMV_TPR");
        forced = e ? atoi(e) : 0;
    }
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16 || forced == 32) return forced;
    double avg = (double)P->nnz / (n > 0 ? n : 1);
    int tpr = 2;
    while (tpr < 32 && tpr * 4 < avg) tpr *= 2;  // ~4 entries per thread
    return tpr;
}

__global__ void k_col16(int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                        int16_t* __restrict__ col16, int* __restrict__ overflow) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    bool bad = false;
    for (int t = rowptr[r]; t < rowptr[r + 1]; ++t) {
        const int d = colidx[t] - r;
        if (d < -32768 || d > 32767) bad = true;
        col16[t] = (int16_t)d;
    }
    if (bad) *overflow = 1;
}

// builds (once) the 16-bit column offsets of a pattern; false when some |col - row| does not fit
static bool ensure_col16(fdb_space* s, Pattern* P) {
    if (P->col16_state != 0) return P->col16_state > 0;
    P->col16_state = -1;
    if (getenv("FDB_NO_COL16")) return false;
    DevBuf<int> flag;
    if (flag.alloc(1) != FDB_OK || P->col16.alloc((size_t)P->nnz) != FDB_OK) return false;
    cudaMemsetAsync(flag.p, 0, sizeof(int), s->stream);
    k_col16<<<(s->n_dofs + 255) / 256, 256, 0, s->stream>>>(s->n_dofs, P->rowptr.p, P->colidx.p, P->col16.p, flag.p);
    int h = 1;
    if (cudaMemcpyAsync(&h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) return false;
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) return false;
    if (h == 0) P->col16_state = 1;
    else P->col16.release();
    return P->col16_state > 0;
}

// ---- sliced-ELL SpMV --------------------------------------------------------------------------------------------------
__global__ void k_sell_slice_len(int n, int n_slices, const int32_t* __restrict__ rowptr, int32_t* __restrict__ slots) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_slices) return;
    if (s == n_slices) { slots[s] = 0; return; }
    int m = 0;
    for (int r = 32 * s; r < 32 * s + 32 && r < n; ++r) m = max(m, rowptr[r + 1] - rowptr[r]);
    slots[s] = 32 * m;
}
__global__ void k_sell_fill(int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                            const int32_t* __restrict__ sell_ptr, int32_t* __restrict__ perm, int32_t* __restrict__ col,
                            int16_t* __restrict__ col16) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    int s = r >> 5, l = r & 31;
    if (32 * s >= n) return;
    const int base = sell_ptr[s], len = (sell_ptr[s + 1] - base) >> 5;
    const int t0 = r < n ? rowptr[r] : 0, rl = r < n ? rowptr[r + 1] - t0 : 0;
    for (int j = 0; j < len; ++j) {
        const int slot = base + 32 * j + l;
        const bool real = j < rl;
        const int c = real ? colidx[t0 + j] : (r < n ? r : 0);
        perm[slot] = real ? t0 + j : -1;
        if (col) col[slot] = c;
        if (col16) col16[slot] = (int16_t)(r < n ? c - r : 0);
    }
}
__global__ void k_sell_values(int64_t slots, const int32_t* __restrict__ perm, const double* __restrict__ val,
                              double* __restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= slots) return;
    const int p = perm[t];
    out[t] = p >= 0 ? val[p] : 0.0;
}

// y = A x with one thread per row over the sliced-ELL arrays; NDOT = number of fused dot products:
//   1: part0 += y.w      2: part0 += y.y, part1 += y.w   (BiCGSTAB's t.t and t.s)
template <int NDOT, bool C16>
__global__ void __launch_bounds__(VB)
k_spmv_sell(int n, const int32_t* __restrict__ sell_ptr, const void* __restrict__ cols, const double* __restrict__ val,
            const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ w,
            double* __restrict__ part0, double* __restrict__ part1, const int* __restrict__ done) {
    __shared__ double sh[VB / 32];
    if (done && *done) return;
    const int lane = threadIdx.x & 31;
    double a0 = 0, a1 = 0;
    const int n_pad = (n + 31) & ~31;
#ifndef FDB_NO_STREAM_HINT
    const unsigned long long pol = l2_evict_first_policy();
#endif
    for (int row = blockIdx.x * VB + threadIdx.x; row < n_pad; row += gridDim.x * VB) {
        const int s = row >> 5;
        const int base = __ldg(sell_ptr + s), len = (__ldg(sell_ptr + s + 1) - base) >> 5;
        double sum = 0;
        if (row < n) {  // rows of the last, partial slice only
#ifndef FDB_NO_STREAM_HINT
            // memory-level parallelism: eight (value, column) pairs are requested at once, then their eight x entries, then
            // the FMAs in slot order (the same sum as a sequential loop)
            for (int j = 0; j < len; j += 8) {
                double a[8], xv[8];
                int c[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool live = j + u < len;
                    const int slot = base + 32 * (j + u) + lane;
                    a[u] = live ? ld_stream(val + slot, pol) : 0.0;
                    if constexpr (C16) c[u] = live ? row + ld_stream(static_cast<const int16_t*>(cols) + slot, pol) : row;
                    else c[u] = live ? ld_stream(static_cast<const int32_t*>(cols) + slot, pol) : row;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) xv[u] = __ldg(x + c[u]);
#pragma unroll
                for (int u = 0; u < 8; ++u) sum += a[u] * xv[u];
            }
#else
            for (int j = 0; j < len; ++j) {
                const int slot = base + 32 * j + lane;
                int c;
                if constexpr (C16) c = row + (int)__ldg(static_cast<const int16_t*>(cols) + slot);
                else c = __ldg(static_cast<const int32_t*>(cols) + slot);
                sum += __ldg(val + slot) * __ldg(x + c);
            }
#endif
        }
        if (row < n) {
            y[row] = sum;
            if (NDOT == 1) a0 += sum * w[row];
            if (NDOT == 2) { a0 += sum * sum; a1 += sum * w[row]; }
        }
    }
    if (NDOT >= 1) {
        a0 = block_sum(a0, sh);
        if (threadIdx.x == 0) part0[blockIdx.x] = a0;
    }
    if (NDOT == 2) {
        a1 = block_sum(a1, sh);
        if (threadIdx.x == 0) part1[blockIdx.x] = a1;
    }
}

// builds (once per pattern) the sliced-ELL index arrays; false when padding would exceed 25 %.
// force: build whatever the padding -- partitioned matrices must take the same solver path on every rank, so the choice
// cannot depend on a local property of the rank's rows.
static bool ensure_sell(fdb_space* s, Pattern* P, bool force = false) {
    if (P->sell_state > 0) return true;
    if (P->sell_state < 0 && !force) return false;
    P->sell_state = -1;
    if (getenv("FDB_NO_SELL") && !force) return false;
    const int n = s->n_dofs, n_slices = (n + 31) / 32;
    cudaStream_t st = s->stream;
    if (P->sell_ptr.alloc((size_t)n_slices + 1) != FDB_OK) return false;
    k_sell_slice_len<<<(n_slices + 256) / 256, 256, 0, st>>>(n, n_slices, P->rowptr.p, P->sell_ptr.p);
    {
        size_t tb = 0;
        if (cub::DeviceScan::ExclusiveSum(nullptr, tb, P->sell_ptr.p, P->sell_ptr.p, n_slices + 1, st) != cudaSuccess) return false;
        DevBuf<char> tmp;
        if (tmp.alloc(tb) != FDB_OK) return false;
        if (cub::DeviceScan::ExclusiveSum(tmp.p, tb, P->sell_ptr.p, P->sell_ptr.p, n_slices + 1, st) != cudaSuccess) return false;
        if (cudaStreamSynchronize(st) != cudaSuccess) return false;
    }
    int32_t slots = 0;
    if (cudaMemcpy(&slots, P->sell_ptr.p + n_slices, sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
    if (slots <= 0 || (!force && (double)slots > 1.25 * (double)P->nnz + 1024.0)) { P->sell_ptr.release(); return false; }
    const bool c16 = ensure_col16(s, P);
    if (P->sell_perm.alloc(slots) != FDB_OK) return false;
    if (c16) { if (P->sell_col16.alloc(slots) != FDB_OK) return false; }
    else if (P->sell_col.alloc(slots) != FDB_OK) return false;
    k_sell_fill<<<(32 * n_slices + 255) / 256, 256, 0, st>>>(n, P->rowptr.p, P->colidx.p, P->sell_ptr.p, P->sell_perm.p,
                                                          c16 ? nullptr : P->sell_col.p, c16 ? P->sell_col16.p : nullptr);
    if (cudaStreamSynchronize(st) != cudaSuccess) return false;
    P->sell_slots = slots;
    P->sell_state = 1;
    return true;
}

// refreshes the matrix values in sliced-ELL order when they have changed since the last SpMV
static int sell_values(fdb_matrix* A) {
    const Pattern* P = A->pat;
    if (A->sell_version == A->val_version && A->sell_val.n >= (size_t)P->sell_slots) return FDB_OK;
    if (A->sell_val.n < (size_t)P->sell_slots) FDB_TRY(A->sell_val.alloc((size_t)P->sell_slots));
    k_sell_values<<<(unsigned)((P->sell_slots + 255) / 256), 256, 0, A->space->stream>>>(P->sell_slots, P->sell_perm.p,
                                                                                      A->val.p, A->sell_val.p);
    FDB_CUDA(cudaGetLastError());
    A->sell_version = A->val_version;
    return FDB_OK;
}

bool sell_view_ready(fdb_matrix* A) {
    if (!ensure_sell(A->space, const_cast<Pattern*>(A->pat), A->part != nullptr)) return false;
    return sell_values(A) == FDB_OK;
}

template <bool DOT>
static int launch_spmv(fdb_matrix* A, int grid, const double* x, double* y, const double* w, double* part,
                       const int* done) {
    fdb_space* s = A->space;
    const Pattern* P = A->pat;
    const int n = A->part ? A->part->n_owned : s->n_dofs;  // rows computed by this rank
    // partitioned matrices use the sliced-ELL view too (slices of the local rows; halo columns index the tail of x)
    if (ensure_sell(s, const_cast<Pattern*>(P), A->part != nullptr)) {
        FDB_TRY(sell_values(A));
        if (P->col16_state > 0)
            k_spmv_sell<DOT ? 1 : 0, true><<<grid, VB, 0, s->stream>>>(n, P->sell_ptr.p, P->sell_col16.p, A->sell_val.p, x, y,
                                                                      w, part, nullptr, done);
        else
            k_spmv_sell<DOT ? 1 : 0, false><<<grid, VB, 0, s->stream>>>(n, P->sell_ptr.p, P->sell_col.p, A->sell_val.p, x, y,
                                                                       w, part, nullptr, done);
        FDB_CUDA(cudaGetLastError());
        return FDB_OK;
    }
    const bool c16 = ensure_col16(s, const_cast<Pattern*>(P)) && !A->part;
    const void* cols = c16 ? static_cast<const void*>(P->col16.p) : static_cast<const void*>(P->colidx.p);
#define FDB_SPMV(T)                                                                                              \
    do {                                                                                                         \
        if (c16) k_spmv<T, DOT, true><<<grid, VB, 0, s->stream>>>(n, P->rowptr.p, cols, A->val.p, x, y, w, part, done); \
        else k_spmv<T, DOT, false><<<grid, VB, 0, s->stream>>>(n, P->rowptr.p, cols, A->val.p, x, y, w, part, done);    \
    } while (0)
    switch (pick_tpr(P, n)) {
    case 1: FDB_SPMV(1); break;
    case 2: FDB_SPMV(2); break;
    case 4: FDB_SPMV(4); break;
    case 8: FDB_SPMV(8); break;
    case 16: FDB_SPMV(16); break;
    default: FDB_SPMV(32); break;
    }
#undef FDB_SPMV
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

static int solver_grid(const fdb_space* s) { return s->sm_count * 8; }

int spmv(fdb_matrix* A, const double* x, double* y) {
    FDB_CHECK(A && A->assembled, FDB_ERR_STATE, "matrix has not been assembled");
    if (A->part) FDB_TRY(halo_exchange(A, const_cast<double*>(x)));  // fills the halo tail of x
    return launch_spmv<false>(A, solver_grid(A->space), x, y, nullptr, nullptr, nullptr);
}

// per-rank sums of up to three partial arrays (distributed solve: input of the all-reduce)
__global__ void __launch_bounds__(VB)
k_collapse(int np, const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
           double* __restrict__ out) {
    __shared__ double sh[VB / 32];
    double va = sum_partials(a, np, sh);
    double vb = b ? sum_partials(b, np, sh) : 0.0;
    double vc = c ? sum_partials(c, np, sh) : 0.0;
    if (threadIdx.x == 0) { out[0] = va; out[1] = vb; out[2] = vc; }
}

int solve(fdb_matrix* A, const double* b, double* x, const fdb_solver_opts* o, fdb_solve_stats* stats) {
    FDB_CHECK(A && A->assembled, FDB_ERR_STATE, "solver must be initialized first!");
    FDB_CHECK(o, FDB_ERR_ARG, "null solver options");
    FDB_CHECK(o->kind == FDB_SOLVER_CG || o->kind == FDB_SOLVER_BICGSTAB, FDB_ERR_ARG, "unknown solver kind");
    {   // whole loop in one cooperative sliced-ELL kernel when possible (partitioned matrices with a peer-memory plan)
        bool handled = false;
        int rc = solve_persistent_sell(A, b, x, o, stats, &handled);
        if (handled || rc != FDB_OK) return rc;
    }
    if (o->kind == FDB_SOLVER_CG) {  // first persistent form (CSR-vector SpMV), opt-in
        bool handled = false;
        int rc = solve_cg_persistent(A, b, x, o, stats, &handled);
        if (handled || rc != FDB_OK) return rc;
    }
    fdb_space* s = A->space;
    const Pattern* P = A->pat;
    Partition* part = A->part;
    cudaStream_t st = s->stream;
    const int ld = s->n_dofs;                      // vector length (owned + halo when partitioned)
    const int n = part ? part->n_owned : ld;       // rows / vector entries this rank iterates on
    const int G = solver_grid(s);
    const int np = G;
    const long long n_glob = part ? part->n_global : (long long)ld;   // identical on every rank
    const int maxit = o->maxit > 0 ? o->maxit : (int)std::min<long long>(10 * std::max<long long>(n_glob, 1), 2000000000LL);
    const int every = o->check_every > 0 ? o->check_every : 32;
    const bool jac = o->jacobi != 0;

    // workspace: 9 vectors, partial arrays, residual history (ring), scalars
    const size_t nv = 9;
    const int hist_cap = 1 << 16;
    if (A->work.n < nv * (size_t)ld) FDB_TRY(A->work.alloc(nv * (size_t)ld));
    if (A->partials.n < 8 * (size_t)np + 64) FDB_TRY(A->partials.alloc(8 * (size_t)np + 64));
    if (A->hist.n < (size_t)hist_cap) FDB_TRY(A->hist.alloc((size_t)hist_cap));
    double* W = A->work.p;
    double *r = W, *p = W + (size_t)ld, *q = W + 2 * (size_t)ld, *z = W + 3 * (size_t)ld, *dinv = W + 4 * (size_t)ld;
    double *r0 = W + 5 * (size_t)ld, *sv = W + 6 * (size_t)ld, *tv = W + 7 * (size_t)ld, *zs2 = W + 8 * (size_t)ld;
    double* PA = A->partials.p;
    double *part0 = PA, *part1 = PA + np, *part2 = PA + 2 * np, *part3 = PA + 3 * np, *part4 = PA + 4 * np,
           *part_bb = PA + 5 * np;
    Scal* sc = reinterpret_cast<Scal*>(PA + 8 * (size_t)np);
    int* done = &sc->done;

    // Distributed mode: after a kernel has produced per-block partials, they are collapsed to one value per rank,
    // all-reduced, and the consumers read the global value (a "partial array" of length 1).  Slots are double
    // buffered by iteration parity so a value is never overwritten while a late block of its consumer still reads it.
    double* L = part ? part->stage.p : nullptr;       // [2][8] per-rank sums
    double* GS = part ? part->stage.p + 16 : nullptr; // [2][8] global sums
    auto reduce3 = [&](int par, int slot, const double* a, const double* b2, const double* c, int count) -> int {
        k_collapse<<<1, VB, 0, st>>>(np, a, b2, c, L + par * 8 + slot);
        FDB_CUDA(cudaGetLastError());
        return allreduce_sum(A, L + par * 8 + slot, GS + par * 8 + slot, count);
    };

    if (jac) {
        k_jacobi<<<(n + 255) / 256, 256, 0, st>>>(n, P->diag.p, A->val.p, dinv);
        FDB_CUDA(cudaGetLastError());
    }
    const double* dv = jac ? dinv : nullptr;

    cudaEvent_t ev0, ev1;
    FDB_CUDA(cudaEventCreate(&ev0));
    FDB_CUDA(cudaEventCreate(&ev1));
    FDB_CUDA(cudaEventRecord(ev0, st));

    Scal h;
    int launched = 0;
    if (o->kind == FDB_SOLVER_CG) {
        // single GPU: part0/part1 = rz (ping-pong), part2 = pq, part3 = rr.   distributed slots: 0 pq, 1 rz, 2 rr, 3 bb
        if (part) FDB_TRY(halo_exchange(A, x));
        FDB_TRY(launch_spmv<false>(A, G, x, q, nullptr, nullptr, nullptr));
        double* zz = jac ? z : r;  // without preconditioner z aliases r
        k_cg_init<<<G, VB, 0, st>>>(n, b, q, dv, r, z, p, part0, part3, part_bb);
        FDB_CUDA(cudaGetLastError());
        if (part) {
            FDB_TRY(reduce3(1, 1, part0, part3, part_bb, 3));
            k_set_threshold<<<1, VB, 0, st>>>(1, GS + 8 + 2, GS + 8 + 3, o->rtol, sc);
        } else {
            k_set_threshold<<<1, VB, 0, st>>>(np, part3, part_bb, o->rtol, sc);
        }
        FDB_CUDA(cudaGetLastError());
        // (An L2 persisting access-policy window over the Krylov vectors was measured slower: C4 CG 77.2 vs 71.4 us per
        // iteration, C3 BiCGSTAB 380 vs 327 -- the vectors already stay in L2 under the evict-first matrix stream.)
        // (Fusing the direction update into the SpMV -- every gathered entry recomputing z[c] + beta p[c] -- was measured
        // slower on C4: 75.7 against 72.2 us per iteration; the second gather costs more than the direction kernel.)
        // one CG iteration as three launches; `par` selects which rz partial array is old / new
        auto launch_iteration = [&](int par) -> int {
            double* rz_old = par ? part1 : part0;
            double* rz_new = par ? part0 : part1;
            if (part) FDB_TRY(halo_exchange(A, p));
            FDB_TRY(launch_spmv<true>(A, G, p, q, p, part2, done));
            if (part) {
                FDB_TRY(reduce3(par, 0, part2, nullptr, nullptr, 1));
                k_cg_update<<<G, VB, 0, st>>>(n, 1, GS + par * 8 + 0, GS + (par ^ 1) * 8 + 1, p, q, dv, x, r, z, rz_new, part3, sc);
                FDB_TRY(reduce3(par, 1, rz_new, part3, nullptr, 2));
                k_cg_direction<<<G, VB, 0, st>>>(n, 1, GS + par * 8 + 1, GS + (par ^ 1) * 8 + 1, GS + par * 8 + 2, zz, p, sc,
                                                 A->hist.p, maxit, hist_cap);
            } else {
                k_cg_update<<<G, VB, 0, st>>>(n, np, part2, rz_old, p, q, dv, x, r, z, rz_new, part3, sc);
                k_cg_direction<<<G, VB, 0, st>>>(n, np, rz_new, rz_old, part3, zz, p, sc, A->hist.p, maxit, hist_cap);
            }
            return FDB_OK;
        };
        // Single GPU: two iterations (even + odd parity) are captured once into a CUDA graph and replayed; the
        // iteration counter, the convergence flag and the iteration budget all live on the device, so a replay after
        // convergence is a sequence of no-ops.  (NCCL calls of the partitioned path are launched directly.)
        cudaGraphExec_t gexec = nullptr;
        if (!part && !getenv("FDB_NO_GRAPH")) {
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                int rc0 = launch_iteration(0), rc1 = launch_iteration(1);
                cudaError_t ec = cudaStreamEndCapture(st, &graph);
                if (rc0 == FDB_OK && rc1 == FDB_OK && ec == cudaSuccess && graph &&
                    cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess)
                    gexec = nullptr;
                if (graph) cudaGraphDestroy(graph);
            }
            cudaGetLastError();  // a failed capture must not poison the direct-launch fallback
        }
        while (launched < maxit) {
            int stop = launched + every < maxit ? launched + every : maxit;
            if (gexec) {
                stop = launched + 2 * ((stop - launched + 1) / 2);
                for (int it = launched; it < stop; it += 2) FDB_CUDA(cudaGraphLaunch(gexec, st));
            } else {
                for (int it = launched; it < stop; ++it) FDB_TRY(launch_iteration(it & 1));
            }
            FDB_CUDA(cudaGetLastError());
            launched = stop;
            FDB_CUDA(cudaMemcpyAsync(&h, sc, sizeof(Scal), cudaMemcpyDeviceToHost, st));
            FDB_CUDA(cudaStreamSynchronize(st));
            if (h.done) break;
        }
        if (gexec) cudaGraphExecDestroy(gexec);
    } else {
        // single GPU: part0 = rho = r0.r, part1 = r0.v, part2 = tt, part3 = ts, part4 = rr.
        // distributed slots: 0 rho, 1 rr, 2 r0v, 3 tt, 4 ts, 5 bb  (rho/rr are produced together, tt/ts together)
        double* vv = q;
        double* y = jac ? z : p;      // y = M^-1 p (aliases p without preconditioner)
        double* zs = jac ? zs2 : sv;  // z = M^-1 s (aliases s without preconditioner)
        const int tpr = pick_tpr(P, n);
        const bool c16 = ensure_col16(s, const_cast<Pattern*>(P)) && !part;
        const void* cols = c16 ? static_cast<const void*>(P->col16.p) : static_cast<const void*>(P->colidx.p);
        // On a breakdown (rho or r0.v vanish) nothing is updated and the iteration restarts from the current x with
        // a fresh shadow residual, as Eigen's BiCGSTAB does.
        for (int restarts = 0;; ++restarts) {
            if (part) FDB_TRY(halo_exchange(A, x));
            FDB_TRY(launch_spmv<false>(A, G, x, q, nullptr, nullptr, nullptr));
            k_bi_init<<<G, VB, 0, st>>>(n, b, q, r, r0, p, vv, part0, part4, part_bb);
            FDB_CUDA(cudaGetLastError());
            if (part) {
                FDB_TRY(reduce3(1, 0, part0, part4, nullptr, 2));
                FDB_TRY(reduce3(1, 5, part_bb, nullptr, nullptr, 1));
                k_set_threshold<<<1, VB, 0, st>>>(1, GS + 8 + 1, GS + 8 + 5, o->rtol, sc, restarts > 0);
            } else {
                k_set_threshold<<<1, VB, 0, st>>>(np, part4, part_bb, o->rtol, sc, restarts > 0);
            }
            FDB_CUDA(cudaGetLastError());
            const int it_start = launched;
            while (launched < maxit) {
                int stop = launched + every < maxit ? launched + every : maxit;
                for (int it = launched; it < stop; ++it) {
                    const int k = it - it_start, par = k & 1;
                    const double* g_rho = part ? GS + (par ^ 1) * 8 + 0 : part0;  // produced by the previous iteration
                    const int npi = part ? 1 : np;
                    k_bi_p<<<G, VB, 0, st>>>(n, npi, k == 0, g_rho, r, vv, dv, p, y, sc);
                    k_bi_store_rho<<<1, VB, 0, st>>>(npi, g_rho, sc);
                    if (part) FDB_TRY(halo_exchange(A, y));
                    FDB_TRY(launch_spmv<true>(A, G, y, vv, r0, part1, done));
                    if (part) FDB_TRY(reduce3(par, 2, part1, nullptr, nullptr, 1));
                    k_bi_s<<<G, VB, 0, st>>>(n, npi, part ? GS + par * 8 + 2 : part1, r, vv, dv, sv, zs, sc);
                    if (part) FDB_TRY(halo_exchange(A, zs));
                    if (P->sell_state > 0) {
                        FDB_TRY(sell_values(A));
                        if (P->col16_state > 0)
                            k_spmv_sell<2, true><<<G, VB, 0, st>>>(n, P->sell_ptr.p, P->sell_col16.p, A->sell_val.p, zs, tv, sv,
                                                                   part2, part3, done);
                        else
                            k_spmv_sell<2, false><<<G, VB, 0, st>>>(n, P->sell_ptr.p, P->sell_col.p, A->sell_val.p, zs, tv, sv,
                                                                    part2, part3, done);
                    } else {
#define FDB_TT(T)                                                                                                      \
    do {                                                                                                               \
        if (c16) k_spmv_tt_ts<T, true><<<G, VB, 0, st>>>(n, P->rowptr.p, cols, A->val.p, zs, tv, sv, part2, part3, sc);  \
        else k_spmv_tt_ts<T, false><<<G, VB, 0, st>>>(n, P->rowptr.p, cols, A->val.p, zs, tv, sv, part2, part3, sc);     \
    } while (0)
                    switch (tpr) {
                    case 1: FDB_TT(1); break;
                    case 2: FDB_TT(2); break;
                    case 4: FDB_TT(4); break;
                    case 8: FDB_TT(8); break;
                    case 16: FDB_TT(16); break;
                    default: FDB_TT(32); break;
                    }
#undef FDB_TT
                    }
                    if (part) FDB_TRY(reduce3(par, 3, part2, part3, nullptr, 2));
                    k_bi_x<<<G, VB, 0, st>>>(n, npi, part ? GS + par * 8 + 3 : part2, part ? GS + par * 8 + 4 : part3, y, zs,
                                             sv, tv, r0, x, r, part4, part0, sc);
                    if (part) FDB_TRY(reduce3(par, 0, part0, part4, nullptr, 2));
                    k_bi_finish<<<1, VB, 0, st>>>(npi, part ? GS + par * 8 + 1 : part4, part ? GS + par * 8 + 3 : part2,
                                                  part ? GS + par * 8 + 4 : part3, sc, A->hist.p, it, hist_cap);
                }
                FDB_CUDA(cudaGetLastError());
                launched = stop;
                FDB_CUDA(cudaMemcpyAsync(&h, sc, sizeof(Scal), cudaMemcpyDeviceToHost, st));
                FDB_CUDA(cudaStreamSynchronize(st));
                if (h.done) break;
            }
            const bool broke = h.done && h.breakdown == 1 && !(h.rr <= h.thr);
            if (!broke || restarts >= 16 || h.iters >= maxit) break;
            launched = h.iters;  // iterations after the breakdown were no-ops
        }
    }
    FDB_CUDA(cudaEventRecord(ev1, st));
    FDB_CUDA(cudaMemcpyAsync(&h, sc, sizeof(Scal), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (h.breakdown == 2) {  // zero right-hand side
        FDB_CUDA(cudaMemsetAsync(x, 0, sizeof(double) * n, st));
        FDB_CUDA(cudaStreamSynchronize(st));
        h.rr = 0;
    }
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
