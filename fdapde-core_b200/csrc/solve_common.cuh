// Helpers shared by the solver translation units (solve.cu, solve_persistent.cu).
#pragma once
#include "common.cuh"

namespace fdb {

constexpr int VB = 256;  // block size of every solver kernel (fixed: the partial-sum order depends on it)

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

// device-side scalar block
struct Scal {
    double bb, thr, rho, alpha, omega, rr;
    int done, iters, breakdown, pad;
    double r0sq;  // |r0|^2 of the current (re)start (BiCGSTAB breakdown test)
};

// One 128-byte line per (rank, reduction point): three sums + a sequence number written last.
struct alignas(128) RedLine {
    double v[3];
    unsigned long long seq;
    double pad[12];
};

// device-visible description of the peer-memory plan (all pointers are valid in THIS process)
struct PeerView {
    int world, rank, n_nbr;
    int n_owned;
    const int32_t* send_idx;           // owned local indices to push, grouped by neighbour
    int send_off[9];                   // prefix offsets per neighbour (<= 8 neighbours)
    double* nbr_p_halo[8];             // where neighbour i expects my entries inside ITS p buffer 0 (peer memory)
    long long nbr_ld[8];               // neighbour i's vector length: its p buffer 1 starts nbr_ld doubles further
    unsigned long long* nbr_flag[8];   // neighbour i's "halo from me has arrived" flag (peer memory)
    unsigned long long* my_flag;       // my flags, one per neighbour slot (local memory, written by peers)
    RedLine* red_of[8];                // reduction buffer of every rank: [4 points][world] lines (peer memory)
    RedLine* my_red;                   // == red_of[rank]
    int* error;                        // set when a wait times out
    unsigned long long seq0;           // first sequence number of this solve (epoch << 32)
};

// layout of a rank's exported exchange buffer (bytes), as every peer computes it from that rank's vector length
struct PeerLayout {
    size_t off_flags, off_red, off_error, bytes;
    static PeerLayout of(size_t ld, int world) {
        PeerLayout L;
        size_t vec = 5 * ld * sizeof(double);  // [r | p0 | p1 | q | z]
        L.off_flags = (vec + 127) / 128 * 128;
        L.off_red = L.off_flags + 128;                       // 8 flags of 8 bytes, padded to one line... (16 x 8 B)
        L.off_error = L.off_red + (size_t)(4 * world + 2) * sizeof(RedLine);
        L.bytes = L.off_error + 128;
        return L;
    }
};

// solve_persistent.cu: whole CG loop in one cooperative kernel (single GPU, or multi-GPU over peer memory)
int solve_cg_persistent(fdb_matrix* A, const double* b, double* x, const fdb_solver_opts* o, fdb_solve_stats* stats,
                        bool* handled);
int pick_tpr(const Pattern* P, int n);
int& persistent_mode();

}  // namespace fdb
