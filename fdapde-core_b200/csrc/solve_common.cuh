// Helpers shared by the solver translation units (solve.cu, solve_persistent.cu).
#pragma once
#include "common.cuh"

namespace fdb {

constexpr int VB = 256;  // block size of every solver kernel (fixed: the partial-sum order depends on it)

// Matrix streams (values, columns) are read exactly once per SpMV: loads carry an L2 evict-first policy and skip L1, so the
// 126 MB L2 keeps the Krylov vectors (a few tens of MB, touched several times per iteration) instead of matrix lines.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double ld_stream(const double* p, unsigned long long pol) {
    double v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int ld_stream(const int32_t* p, unsigned long long pol) {
    int v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int ld_stream(const int16_t* p, unsigned long long pol) {
    short v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.s16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
    return (int)v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block sum, result broadcast to all threads
__device__ __forceinline__ double block_sum(double v, double* sh /* >= 8 doubles */) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0;
#pragma unroll
    for (int k = 0; k < VB / 32; ++k) t += sh[k];
    return t;
}

// sum of `np` per-block partials, identical in every block
__device__ __forceinline__ double sum_partials(const double* __restrict__ part, int np, double* sh) {
    double v = 0;
    for (int k = threadIdx.x; k < np; k += VB) v += part[k];
    return block_sum(v, sh);
}

// the same sums for up to three partial arrays at once: one pass of loads and one pair of barriers instead of three
// (each value is accumulated and reduced in exactly the order of sum_partials, so the results are bit-identical to it)
__device__ __forceinline__ void sum_partials3(const double* __restrict__ a, const double* __restrict__ b,
                                              const double* __restrict__ c, int np, double* sh /* 3 * VB / 32 doubles */,
                                              double& va, double& vb, double& vc) {
    double x = 0, y = 0, z = 0;
    for (int k = threadIdx.x; k < np; k += VB) {
        x += a[k];
        if (b) y += b[k];
        if (c) z += c[k];
    }
    x = warp_sum(x); y = warp_sum(y); z = warp_sum(z);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { sh[w] = x; sh[VB / 32 + w] = y; sh[2 * (VB / 32) + w] = z; }
    __syncthreads();
    va = 0; vb = 0; vc = 0;
#pragma unroll
    for (int k = 0; k < VB / 32; ++k) { va += sh[k]; vb += sh[VB / 32 + k]; vc += sh[2 * (VB / 32) + k]; }
}

// device-side scalar block
struct Scal {
    double bb, thr, rho, alpha, omega, rr;
    int done, iters, breakdown, pad;
    double r0sq;  // |r0|^2 of the current (re)start (BiCGSTAB breakdown test)
};

// Peer-memory exchange of the persistent multi-GPU CG uses a flag-in-data ("LL") encoding: every fp64 value travels as
// two 64-bit words {low 32 bits | tag << 32, high 32 bits | tag << 32}.  A 64-bit store is atomic, so a receiver that
// reads both words with the expected tag has the complete value -- no fences, no separate flags, no acquire/release:
// the consumer simply spins on the data it needs, when it needs it.  Tags increase monotonically over the life of a
// matrix (all ranks count the same exchanges), buffers start zeroed, tag 0 is never used.
// (struct LLWord is declared in common.cuh)

// device-visible description of the peer-memory plan (all pointers are valid in THIS process)
struct PeerView {
    int world, rank, n_nbr;
    int n_owned, n_halo;
    const int32_t* send_idx;           // owned local indices to push, grouped by neighbour
    int send_off[9];                   // prefix offsets per neighbour (<= 8 neighbours)
    LLWord* nbr_halo[8];               // where neighbour i receives my entries: its halo buffer 0 + its offset for me
    long long nbr_n_halo[8];           // neighbour i's halo size (its buffer 1 starts that many LLWords further)
    LLWord* my_halo;                   // my halo receive buffers [2][n_halo]
    LLWord* red_of[8];                 // reduction buffers of every rank: [4 points][world][3 values]
    LLWord* my_red;                    // == red_of[rank]
    int* error;                        // set when a wait times out
    unsigned tag0;                     // last tag used by the previous solve on this matrix
};

// layout of a rank's exported exchange buffer (bytes), as every peer computes it from that rank's sizes
struct PeerLayout {
    size_t off_halo, off_red, off_error, bytes;
    static PeerLayout of(size_t n_halo, int world) {
        PeerLayout L;
        L.off_halo = 0;
        L.off_red = (2 * n_halo * sizeof(LLWord) + 127) / 128 * 128;
        L.off_error = L.off_red + (size_t)(4 * world * 4) * sizeof(LLWord);   // [4 points][world][up to 4 values]
        L.off_error = (L.off_error + 127) / 128 * 128;
        L.bytes = L.off_error + 128;
        return L;
    }
};

// solve_persistent.cu: whole CG loop in one cooperative kernel (single GPU, or multi-GPU over peer memory)
int solve_cg_persistent(fdb_matrix* A, const double* b, double* x, const fdb_solver_opts* o, fdb_solve_stats* stats,
                        bool* handled);
// solve_peer.cu: single-reduction CG and BiCGSTAB as one persistent sliced-ELL kernel per GPU
int solve_persistent_sell(fdb_matrix* A, const double* b, double* x, const fdb_solver_opts* o, fdb_solve_stats* stats,
                          bool* handled);
int& peer_sell_mode();
bool sell_view_ready(fdb_matrix* A);   // solve.cu: builds / refreshes the sliced-ELL arrays; false when not worth it
int pick_tpr(const Pattern* P, int n);
int& persistent_mode();

}  // namespace fdb
