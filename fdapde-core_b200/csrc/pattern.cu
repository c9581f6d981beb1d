// K2: sparsity pattern + scatter map, built on the device by sort/unique.
//
// Structural restatement of what the reference obtains from
//   triplet emission                fem_assembler.h:79-110  (cells ascending, i outer, j inner, `>=` filter :96)
//   setFromTriplets/makeCompressed  fem_assembler.h:112-113 (duplicates summed in emission order, zeros kept)
//   selfadjointView<Lower>()        fem_assembler.h:116-117 (lower triangle mirrored to a full matrix)
// Every emitted triplet gets a slot in a contribution list sorted by (row, col, emission order); one segment of
// that list is one stored entry, and summing a segment left to right reproduces Eigen's duplicate order.
// The sort is a stable LSD radix sort (CUB), so equal keys keep ascending (cell, slot) order.
#include <cub/cub.cuh>

#include "common.cuh"

namespace fdb {

static inline int bits_for(int64_t n) {  // smallest b with n <= 2^b
    int b = 1;
    while ((int64_t(1) << b) < n) ++b;
    return b;
}

__host__ __device__ inline int sym_pair_count(int nb) { return nb * (nb + 1) / 2; }

// emission slots of one cell: symmetric => local pairs a <= b (row = larger dof), else all (i, j)
__global__ void k_emit_keys(int n_cells, int nb, int ne, int symmetric, int shift, const int32_t* __restrict__ dofs,
                            uint64_t* __restrict__ keys, uint32_t* __restrict__ ids) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = (int64_t)n_cells * ne;
    if (t >= total) return;
    int e = (int)(t / ne), s = (int)(t % ne);
    int i, j;
    if (symmetric) {
        // s-th pair of the upper-triangular enumeration (0,0),(0,1),...,(0,nb-1),(1,1),...
        i = 0;
        int rem = s;
        while (rem >= nb - i) { rem -= nb - i; ++i; }
        j = i + rem;
    } else {
        i = s / nb;
        j = s % nb;
    }
    uint32_t di = (uint32_t)dofs[(size_t)i * n_cells + e], dj = (uint32_t)dofs[(size_t)j * n_cells + e];
    uint32_t row = di, col = dj;
    if (symmetric && row < col) { uint32_t tmp = row; row = col; col = tmp; }
    keys[t] = ((uint64_t)row << shift) | col;
    ids[t] = (uint32_t)t;
}

__global__ void k_flag_heads(int64_t n, const uint64_t* __restrict__ keys, int32_t* __restrict__ flags) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    flags[t] = (t == 0 || keys[t] != keys[t - 1]) ? 1 : 0;
}

// after the inclusive scan: uid = scan[t]-1.  Records segment starts, unique keys and the scatter map.
__global__ void k_segments(int64_t n, int n_cells, int ne, const uint64_t* __restrict__ keys,
                           const uint32_t* __restrict__ ids, const int32_t* __restrict__ scan,
                           int32_t* __restrict__ seg, uint64_t* __restrict__ ukeys, int32_t* __restrict__ pos) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    int32_t u = scan[t] - 1;
    bool head = (t == 0) || (scan[t - 1] != scan[t]);
    if (head) {
        seg[u] = (int32_t)t;
        ukeys[u] = keys[t];
    }
    uint32_t id = ids[t];
    int e = (int)(id / (uint32_t)ne), s = (int)(id % (uint32_t)ne);
    pos[(size_t)s * n_cells + e] = (int32_t)t;
    if (t == n - 1) seg[u + 1] = (int32_t)n;
}

__global__ void k_mirror_keys(int64_t nu, int shift, const uint64_t* __restrict__ ukeys, uint64_t* __restrict__ k2,
                              uint32_t* __restrict__ v2) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= nu) return;
    uint64_t mask = (uint64_t(1) << shift) - 1;
    uint64_t key = ukeys[u];
    uint64_t r = key >> shift, c = key & mask;
    k2[2 * u] = key;
    v2[2 * u] = (uint32_t)(2 * u);
    k2[2 * u + 1] = (r == c) ? (uint64_t(1) << (2 * shift)) : ((c << shift) | r);  // diagonal: sentinel, sorts last
    v2[2 * u + 1] = (uint32_t)(2 * u + 1);
}

__global__ void k_full_from_sorted(int64_t nnz, int shift, const uint64_t* __restrict__ k2,
                                   const uint32_t* __restrict__ v2, int32_t* __restrict__ colidx,
                                   int32_t* __restrict__ dst_a, int32_t* __restrict__ dst_b) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nnz) return;
    uint64_t mask = (uint64_t(1) << shift) - 1;
    colidx[t] = (int32_t)(k2[t] & mask);
    uint32_t id = v2[t];
    if (id & 1u) dst_b[id >> 1] = (int32_t)t;
    else dst_a[id >> 1] = (int32_t)t;
}

__global__ void k_cols_identity(int64_t nnz, int shift, const uint64_t* __restrict__ ukeys,
                                int32_t* __restrict__ colidx, int32_t* __restrict__ dst_a) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nnz) return;
    uint64_t mask = (uint64_t(1) << shift) - 1;
    colidx[t] = (int32_t)(ukeys[t] & mask);
    dst_a[t] = (int32_t)t;
}

// rowptr[r] = first position whose key >= (r << shift)
__global__ void k_rowptr(int n_rows, int64_t nnz, int shift, const uint64_t* __restrict__ keys,
                         int32_t* __restrict__ rowptr) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    uint64_t target = (uint64_t)r << shift;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    rowptr[r] = (int32_t)lo;
}

__global__ void k_diag(int n_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                       int32_t* __restrict__ diag) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    int lo = rowptr[r], hi = rowptr[r + 1];
    int d = -1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        int c = colidx[mid];
        if (c == r) { d = mid; break; }
        if (c < r) lo = mid + 1;
        else hi = mid;
    }
    diag[r] = d;
}

__global__ void k_transpose_perm(int n_rows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                 int32_t* __restrict__ tperm) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    for (int t = rowptr[r]; t < rowptr[r + 1]; ++t) {
        int c = colidx[t];
        int lo = rowptr[c], hi = rowptr[c + 1], f = -1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            int cc = colidx[mid];
            if (cc == r) { f = mid; break; }
            if (cc < r) lo = mid + 1;
            else hi = mid;
        }
        tperm[t] = f;
    }
}

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

template <typename K, typename V>
static int radix_sort_pairs(DevBuf<K>& k_in, DevBuf<K>& k_out, DevBuf<V>& v_in, DevBuf<V>& v_out, int64_t n,
                            int end_bit, cudaStream_t st) {
    FDB_CHECK(n < (int64_t(1) << 31), FDB_ERR_UNSUPPORTED, "more than 2^31 entries in one sort");
    size_t tmp_bytes = 0;
    FDB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, (int)n, 0, end_bit,
                                             st));
    DevBuf<char> tmp;
    FDB_TRY(tmp.alloc(tmp_bytes));
    FDB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, (int)n, 0, end_bit,
                                             st));
    FDB_CUDA(cudaStreamSynchronize(st));
    return FDB_OK;
}

int build_pattern(fdb_space* s, int symmetric) {
    Pattern& P = s->pat[symmetric ? 1 : 0];
    if (P.built) return FDB_OK;
    cudaStream_t st = s->stream;
    const int nb = s->nb, n_cells = s->n_cells, n = s->n_dofs;
    P.symmetric = symmetric != 0;
    P.ne = symmetric ? sym_pair_count(nb) : nb * nb;
    P.n_contrib = (int64_t)n_cells * P.ne;
    FDB_CHECK(P.n_contrib > 0, FDB_ERR_ARG, "empty mesh");
    FDB_CHECK(P.n_contrib < (int64_t(1) << 31), FDB_ERR_UNSUPPORTED,
              "n_cells * entries-per-cell exceeds int32 (Eigen's StorageIndex); partition the mesh across GPUs");
    const int shift = bits_for(n);
    const int64_t nc = P.n_contrib;
    const int B = 256;

    DevBuf<uint64_t> k0, k1, ukeys;
    DevBuf<uint32_t> v0, v1;
    DevBuf<int32_t> scan;
    FDB_TRY(k0.alloc(nc)); FDB_TRY(k1.alloc(nc)); FDB_TRY(v0.alloc(nc)); FDB_TRY(v1.alloc(nc));
    k_emit_keys<<<grid_for(nc, B), B, 0, st>>>(n_cells, nb, P.ne, symmetric, shift, s->dofs.p, k0.p, v0.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(radix_sort_pairs(k0, k1, v0, v1, nc, 2 * shift, st));
    k0.release(); v0.release();

    // heads -> unique ids
    FDB_TRY(scan.alloc(nc));
    k_flag_heads<<<grid_for(nc, B), B, 0, st>>>(nc, k1.p, scan.p);
    FDB_CUDA(cudaGetLastError());
    {
        size_t tmp_bytes = 0;
        FDB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, scan.p, scan.p, (int)nc, st));
        DevBuf<char> tmp;
        FDB_TRY(tmp.alloc(tmp_bytes));
        FDB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, scan.p, scan.p, (int)nc, st));
        FDB_CUDA(cudaStreamSynchronize(st));
    }
    int32_t nu32 = 0;
    FDB_CUDA(cudaMemcpyAsync(&nu32, scan.p + (nc - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    P.n_unique = nu32;
    FDB_TRY(P.seg.alloc(P.n_unique + 1));
    FDB_TRY(ukeys.alloc(P.n_unique));
    FDB_TRY(P.pos.alloc(nc));
    k_segments<<<grid_for(nc, B), B, 0, st>>>(nc, n_cells, P.ne, k1.p, v1.p, scan.p, P.seg.p, ukeys.p, P.pos.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(st));
    k1.release(); v1.release(); scan.release();

    FDB_TRY(P.rowptr.alloc((size_t)n + 1));
    FDB_TRY(P.dst_a.alloc(P.n_unique));
    if (!symmetric) {
        P.nnz = P.n_unique;
        FDB_TRY(P.colidx.alloc(P.nnz));
        k_cols_identity<<<grid_for(P.nnz, B), B, 0, st>>>(P.nnz, shift, ukeys.p, P.colidx.p, P.dst_a.p);
        FDB_CUDA(cudaGetLastError());
        k_rowptr<<<grid_for(n + 1, B), B, 0, st>>>(n, P.nnz, shift, ukeys.p, P.rowptr.p);
        FDB_CUDA(cudaGetLastError());
    } else {
        const int64_t n2 = 2 * P.n_unique;
        DevBuf<uint64_t> m0, m1;
        DevBuf<uint32_t> w0, w1;
        FDB_TRY(m0.alloc(n2)); FDB_TRY(m1.alloc(n2)); FDB_TRY(w0.alloc(n2)); FDB_TRY(w1.alloc(n2));
        k_mirror_keys<<<grid_for(P.n_unique, B), B, 0, st>>>(P.n_unique, shift, ukeys.p, m0.p, w0.p);
        FDB_CUDA(cudaGetLastError());
        FDB_TRY(radix_sort_pairs(m0, m1, w0, w1, n2, 2 * shift + 1, st));
        // number of valid (non-sentinel) keys = rowptr[n]
        k_rowptr<<<grid_for(n + 1, B), B, 0, st>>>(n, n2, shift, m1.p, P.rowptr.p);
        FDB_CUDA(cudaGetLastError());
        int32_t nnz32 = 0;
        FDB_CUDA(cudaMemcpyAsync(&nnz32, P.rowptr.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaStreamSynchronize(st));
        P.nnz = nnz32;
        FDB_TRY(P.colidx.alloc(P.nnz));
        FDB_TRY(P.dst_b.alloc(P.n_unique));
        FDB_CUDA(cudaMemsetAsync(P.dst_b.p, 0xFF, sizeof(int32_t) * P.n_unique, st));
        k_full_from_sorted<<<grid_for(P.nnz, B), B, 0, st>>>(P.nnz, shift, m1.p, w1.p, P.colidx.p, P.dst_a.p, P.dst_b.p);
        FDB_CUDA(cudaGetLastError());
        FDB_CUDA(cudaStreamSynchronize(st));
    }
    FDB_TRY(P.diag.alloc(n));
    k_diag<<<grid_for(n, B), B, 0, st>>>(n, P.rowptr.p, P.colidx.p, P.diag.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(st));
    P.built = true;
    return FDB_OK;
}

int build_transpose_perm(fdb_space* s, Pattern* P) {
    if (P->tperm.p) return FDB_OK;
    FDB_TRY(P->tperm.alloc(P->nnz));
    k_transpose_perm<<<grid_for(s->n_dofs, 128), 128, 0, s->stream>>>(s->n_dofs, P->rowptr.p, P->colidx.p, P->tperm.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    return FDB_OK;
}

// ---- load-vector gather lists (K5): contributions (cell, i) sorted by dof, emission order preserved -----------
__global__ void k_emit_dof_keys(int64_t total, const int32_t* __restrict__ dofs_soa, int n_cells, int nb,
                                uint32_t* __restrict__ keys, uint32_t* __restrict__ ids) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= total) return;
    int e = (int)(t / nb), i = (int)(t % nb);
    keys[t] = (uint32_t)dofs_soa[(size_t)i * n_cells + e];
    ids[t] = (uint32_t)t;
}
__global__ void k_forcing_pos(int64_t total, int n_cells, int nb, const uint32_t* __restrict__ ids,
                              int32_t* __restrict__ pos) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= total) return;
    uint32_t id = ids[t];
    int e = (int)(id / (uint32_t)nb), i = (int)(id % (uint32_t)nb);
    pos[(size_t)i * n_cells + e] = (int32_t)t;
}
__global__ void k_dof_seg(int n_dofs, int64_t total, const uint32_t* __restrict__ keys, int32_t* __restrict__ seg) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d > n_dofs) return;
    int64_t lo = 0, hi = total;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < (uint32_t)d) lo = mid + 1;
        else hi = mid;
    }
    seg[d] = (int32_t)lo;
}

int build_forcing_map(fdb_space* s) {
    ForcingMap& F = s->fmap;
    if (F.built) return FDB_OK;
    cudaStream_t st = s->stream;
    const int64_t total = (int64_t)s->n_cells * s->nb;
    FDB_CHECK(total < (int64_t(1) << 31), FDB_ERR_UNSUPPORTED, "n_cells * n_basis exceeds int32");
    DevBuf<uint32_t> k0, k1, v0, v1;
    FDB_TRY(k0.alloc(total)); FDB_TRY(k1.alloc(total)); FDB_TRY(v0.alloc(total)); FDB_TRY(v1.alloc(total));
    const int B = 256;
    k_emit_dof_keys<<<grid_for(total, B), B, 0, st>>>(total, s->dofs.p, s->n_cells, s->nb, k0.p, v0.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(radix_sort_pairs(k0, k1, v0, v1, total, bits_for(s->n_dofs), st));
    FDB_TRY(F.pos.alloc(total));
    FDB_TRY(F.seg.alloc((size_t)s->n_dofs + 1));
    k_forcing_pos<<<grid_for(total, B), B, 0, st>>>(total, s->n_cells, s->nb, v1.p, F.pos.p);
    FDB_CUDA(cudaGetLastError());
    k_dof_seg<<<grid_for(s->n_dofs + 1, B), B, 0, st>>>(s->n_dofs, total, k1.p, F.seg.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(st));
    F.built = true;
    return FDB_OK;
}

}  // namespace fdb
