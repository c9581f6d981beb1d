// K2: sparsity pattern + scatter map, built on the device by sort/unique.
//
// Structural restatement of what the reference obtains from
//   triplet emission                fem_assembler.h:79-110  (cells ascending, i outer, j inner, `>=` filter :96)
//   setFromTriplets/makeCompressed  fem_assembler.h:112-113 (duplicates summed in emission order, zeros kept)
//   selfadjointView<Lower>()        fem_assembler.h:116-117 (lower triangle mirrored to a full matrix)
// Every emitted triplet gets a slot in a contribution list sorted by (row, col, emission order); one segment of
// that list is one stored entry, and summing a segment left to right reproduces Eigen's duplicate order.
// The sort is a stable LSD radix sort (CUB), so equal keys keep ascending (cell, slot) order.
#include <algorithm>
#include <cub/cub.cuh>

#include "local_matrix.cuh"

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

// ---- fused plan: spatially compact row blocks + their cell lists + shared-memory gather indices -------------------
__device__ __forceinline__ uint64_t spread3(uint64_t x) {  // 21 bits -> every third bit
    x &= 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}

// Morton key of the centroid of one cell incident to each row (the first contribution of the row's last entry)
__global__ void k_row_keys(int n, int M, int ne, int n_cells, int n_nodes, const int32_t* __restrict__ urow,
                           const int32_t* __restrict__ seg, const uint32_t* __restrict__ ids,
                           const int32_t* __restrict__ verts, const double* __restrict__ coords, double lo0, double lo1,
                           double lo2, double sc0, double sc1, double sc2, uint64_t* __restrict__ keys,
                           uint32_t* __restrict__ rows) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    rows[r] = (uint32_t)r;
    if (urow[r + 1] <= urow[r]) { keys[r] = 0; return; }
    int u = urow[r + 1] - 1;
    int e = (int)(ids[seg[u]] / (uint32_t)ne);
    double c[3] = {0, 0, 0};
    for (int k = 0; k <= M; ++k) {
        int v = verts[(size_t)k * n_cells + e];
        for (int d = 0; d < M; ++d) c[d] += coords[(size_t)d * n_nodes + v];
    }
    const double lo[3] = {lo0, lo1, lo2}, sc[3] = {sc0, sc1, sc2};
    uint64_t q[3] = {0, 0, 0};
    for (int d = 0; d < M; ++d) {
        double t = (c[d] / (M + 1) - lo[d]) * sc[d];
        t = t < 0 ? 0 : (t > 2097151.0 ? 2097151.0 : t);
        q[d] = (uint64_t)t;
    }
    keys[r] = spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2);
}

__global__ void k_inverse_perm(int n, const uint32_t* __restrict__ order, int32_t* __restrict__ rorder,
                               int32_t* __restrict__ rank) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    rorder[k] = (int32_t)order[k];
    rank[order[k]] = k;
}

// (block, cell) key of every contribution
__global__ void k_block_cell_keys(int64_t nc, int ne, int shift, int rb, const uint64_t* __restrict__ ukeys,
                                  const uint32_t* __restrict__ ids, const int32_t* __restrict__ scan,
                                  const int32_t* __restrict__ rank, uint64_t* __restrict__ keys) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nc) return;
    uint32_t row = (uint32_t)(ukeys[scan[t] - 1] >> shift);
    uint32_t b = (uint32_t)(rank[row] / rb);
    keys[t] = ((uint64_t)b << 32) | (ids[t] / (uint32_t)ne);
}

__global__ void k_compact_keys(int64_t n, const uint64_t* __restrict__ keys, const int32_t* __restrict__ scan,
                               uint64_t* __restrict__ uniq) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (t == 0 || scan[t] != scan[t - 1]) uniq[scan[t] - 1] = keys[t];
}

__global__ void k_block_ptr(int nblocks, int64_t nu, const uint64_t* __restrict__ uniq, int32_t* __restrict__ ptr) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nblocks) return;
    uint64_t target = (uint64_t)b << 32;
    int64_t lo = 0, hi = nu;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (uniq[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    ptr[b] = (int32_t)lo;
}

__global__ void k_low32(int64_t n, const uint64_t* __restrict__ in, int32_t* __restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = (int32_t)(in[t] & 0xffffffffu);
}

// position of (block, cell) in the ascending list of listed cells
__device__ __forceinline__ int find_pair(const uint64_t* __restrict__ uniq, int lo, int hi, uint64_t key) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (uniq[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// mask of the emission slots of every listed cell that its block actually sums (a cell on the boundary of a block
// contributes only the entries whose row the block owns); OR is order independent => deterministic
__global__ void k_pair_masks(int64_t nc, int ne, int shift, int rb, const uint64_t* __restrict__ ukeys,
                             const uint32_t* __restrict__ ids, const int32_t* __restrict__ scan,
                             const int32_t* __restrict__ rank, const int32_t* __restrict__ bptr,
                             const uint64_t* __restrict__ uniq, unsigned long long* __restrict__ mask) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nc) return;
    uint32_t row = (uint32_t)(ukeys[scan[t] - 1] >> shift);
    int b = rank[row] / rb;
    uint32_t id = ids[t];
    uint32_t e = id / (uint32_t)ne, sl = id % (uint32_t)ne;
    int idx = find_pair(uniq, bptr[b], bptr[b + 1], ((uint64_t)b << 32) | e);
    atomicOr(mask + idx, 1ull << sl);
}

__global__ void k_mask_keys(int64_t total, int ne, const unsigned long long* __restrict__ mask, uint64_t* __restrict__ keys,
                            uint32_t* __restrict__ vals) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    // ascending sort => most needed entries first, identical masks adjacent (ne <= 57: 7 bits of count above the mask)
    const unsigned long long m = mask[i], low = (ne >= 64) ? ~0ull : ((1ull << ne) - 1ull);
    keys[i] = ((unsigned long long)(64 - __popcll(m)) << ne) | (~m & low);
    vals[i] = (uint32_t)i;
}
__global__ void k_block_keys_of(int64_t total, const uint32_t* __restrict__ idx, const uint64_t* __restrict__ uniq,
                                uint64_t* __restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    keys[i] = uniq[idx[i]] >> 32;
}
// final order of the listed cells: cell id, mask, length of the compact record (popcount rounded up to odd: records of
// equal masks then sit at an odd stride, so the stores of a warp are conflict free)
__global__ void k_cell_records(int64_t total, const uint32_t* __restrict__ order, const uint64_t* __restrict__ uniq,
                               const unsigned long long* __restrict__ mask, int32_t* __restrict__ bcells,
                               unsigned long long* __restrict__ bmask, int32_t* __restrict__ len, int32_t* __restrict__ posof) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p > total) return;
    if (p == total) { len[p] = 0; return; }
    uint32_t i = order[p];
    unsigned long long m = mask[i];
    bcells[p] = (int32_t)(uniq[i] & 0xffffffffu);
    bmask[p] = m;
    len[p] = __popcll(m) | 1;
    posof[i] = (int32_t)p;
}
__global__ void k_block_entry_caps(int nblocks, const int32_t* __restrict__ bptr, const int32_t* __restrict__ off,
                                   int32_t* __restrict__ cap) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nblocks) cap[b] = off[bptr[b + 1]] - off[bptr[b]];
}
__global__ void k_cell_bases(int64_t total, const uint64_t* __restrict__ uniq, const uint32_t* __restrict__ order,
                             const int32_t* __restrict__ bptr, const int32_t* __restrict__ off, uint16_t* __restrict__ bbase) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= total) return;
    int b = (int)(uniq[order[p]] >> 32);
    bbase[p] = (uint16_t)(off[p] - off[bptr[b]]);
}

// shared-memory index of every contribution: start of its cell's compact record + rank of its slot among the needed ones
__global__ void k_gather_index(int64_t nc, int ne, int shift, int rb, const uint64_t* __restrict__ ukeys,
                               const uint32_t* __restrict__ ids, const int32_t* __restrict__ scan,
                               const int32_t* __restrict__ rank, const int32_t* __restrict__ bptr,
                               const uint64_t* __restrict__ uniq, const int32_t* __restrict__ posof,
                               const unsigned long long* __restrict__ bmask, const uint16_t* __restrict__ bbase,
                               int lcap /* 0: compact records, else slot-major [slot][lcap] */, uint16_t* __restrict__ lidx) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nc) return;
    uint32_t row = (uint32_t)(ukeys[scan[t] - 1] >> shift);
    int b = rank[row] / rb;
    uint32_t id = ids[t];
    uint32_t e = id / (uint32_t)ne, sl = id % (uint32_t)ne;
    const int idx = find_pair(uniq, bptr[b], bptr[b + 1], ((uint64_t)b << 32) | e);
    if (lcap > 0) { lidx[t] = (uint16_t)(sl * lcap + (idx - bptr[b])); return; }
    const int p = posof[idx];
    lidx[t] = (uint16_t)(bbase[p] + __popcll(bmask[p] & ((1ull << sl) - 1ull)));
}

// bounding box of the nodes (TriangulationBase::range, triangulation.h:52-56) by device reductions
int node_bounding_box(fdb_space* s, double lo[3], double hi[3]) {
    cudaStream_t st = s->stream;
    for (int d = 0; d < 3; ++d) { lo[d] = 0; hi[d] = 1; }
    DevBuf<double> red;
    FDB_TRY(red.alloc(6));
    size_t tb = 0;
    FDB_CUDA(cub::DeviceReduce::Min(nullptr, tb, s->coords.p, red.p, s->n_nodes, st));
    DevBuf<char> tmp;
    FDB_TRY(tmp.alloc(tb));
    for (int d = 0; d < s->N; ++d) {
        FDB_CUDA(cub::DeviceReduce::Min(tmp.p, tb, s->coords.p + (size_t)d * s->n_nodes, red.p + d, s->n_nodes, st));
        FDB_CUDA(cub::DeviceReduce::Max(tmp.p, tb, s->coords.p + (size_t)d * s->n_nodes, red.p + 3 + d, s->n_nodes, st));
    }
    double h[6] = {0, 0, 0, 1, 1, 1};
    FDB_CUDA(cudaMemcpyAsync(h, red.p, sizeof(double) * 6, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    for (int d = 0; d < s->N; ++d) { lo[d] = h[d]; hi[d] = h[3 + d]; }
    return FDB_OK;
}

// rows per block are tried along 512, 256, 128, ... (fine = true: 512, 384, 256, 192, 128, 96, ... for P2 tetrahedra,
// whose blocks are bounded by shared memory)
static inline int next_rb(int rb, bool fine) { return !fine ? rb / 2 : ((rb & (rb - 1)) == 0 ? rb / 4 * 3 : rb / 3 * 2); }

static int build_fused_plan(fdb_space* s, Pattern& P, int shift, const uint64_t* ukeys, const uint32_t* ids,
                            const int32_t* scan, DevBuf<int32_t>& rank, int rb_cap) {
    P.fused = false;
    if (getenv("FDB_NO_FUSED")) return FDB_OK;
    if (s->M != s->N) return FDB_OK;            // manifold cells use the contribution-list path (surface.cu)
    // P2 elements (21 .. 55 entries per cell): a block stores only the entries it sums, as compact records with a 64-bit
    // slot mask per listed cell (non-symmetric P2 tetrahedra have 100 slots: contribution-list path).  P1 elements keep the
    // slot-major array of whole local matrices: at 6 / 10 entries per cell the masks cost more than they save (measured).
    const bool p2tet = s->M == 3 && s->R == 2, compact = s->R == 2, fine = compact;
    if (compact && P.ne > 57) return FDB_OK;
    cudaStream_t st = s->stream;
    const int n = s->n_dofs, B = 256;
    const int64_t nc = P.n_contrib;
    const int smem_limit = 200 * 1024;
    // local entries of one block.  P1 (slot-major, 8 * ne bytes per listed cell), measured on B200: tetrahedra are fastest with
    // 64-row blocks (66 KB, 3 CTAs of 384 threads per SM), triangles with 128-row blocks.  P2 (compact records, 8 bytes per
    // contribution + one pad per listed cell): tetrahedra 34 KB (96 rows, 4 CTAs of 256 threads), triangles 40 KB (256 rows).
    // P1 tetrahedra run the persistent kernel (2 CTAs per SM, 113 KB each): up to 88 KB of local matrices (80-row blocks on
    // a Kuhn mesh) + 32 KB of block lists; measured on C4 (split node copies, 320 threads): 72 rows 0.305, 80 rows 0.301 ms;
    // ensure_fused_plan retries with 8 rows less while two CTAs do not fit
    const bool p1tet = s->M == 3 && s->R == 1;
    // P1 triangles (persistent kernel, 3 CTAs of 320 threads): 256-row blocks, measured 0.080 ms on C2 against 0.088 (128 rows)
    // and 0.084 (512 rows, 2 CTAs)
    int smem_target = p1tet ? 88 * 1024 : (p2tet ? 36 * 1024 : (compact ? 40 * 1024 : 56 * 1024));
    if (const char* e = getenv("FDB_FUSED_SMEM_KB")) smem_target = atoi(e) * 1024;

    FDB_TRY(P.f_urow.alloc((size_t)n + 1));
    k_rowptr<<<grid_for(n + 1, B), B, 0, st>>>(n, P.n_unique, shift, ukeys, P.f_urow.p);
    FDB_CUDA(cudaGetLastError());

    double lo[3], hi[3];
    FDB_TRY(node_bounding_box(s, lo, hi));
    double sc[3];
    for (int d = 0; d < 3; ++d) sc[d] = (hi[d] > lo[d]) ? 2097151.0 / (hi[d] - lo[d]) : 0.0;

    // rows in Morton order of an incident cell
    FDB_TRY(rank.alloc(n));
    FDB_TRY(P.f_rorder.alloc(n));
    {
        DevBuf<uint64_t> rk0, rk1;
        DevBuf<uint32_t> rr0, rr1;
        FDB_TRY(rk0.alloc(n)); FDB_TRY(rk1.alloc(n)); FDB_TRY(rr0.alloc(n)); FDB_TRY(rr1.alloc(n));
        k_row_keys<<<grid_for(n, B), B, 0, st>>>(n, s->M, P.ne, s->n_cells, s->n_nodes, P.f_urow.p, P.seg.p, ids,
                                                s->verts_p, s->coords.p, lo[0], lo[1], lo[2], sc[0], sc[1], sc[2],
                                                rk0.p, rr0.p);
        FDB_CUDA(cudaGetLastError());
        FDB_TRY(radix_sort_pairs(rk0, rk1, rr0, rr1, n, 63, st));
        k_inverse_perm<<<grid_for(n, B), B, 0, st>>>(n, rr1.p, P.f_rorder.p, rank.p);
        FDB_CUDA(cudaGetLastError());
        FDB_CUDA(cudaStreamSynchronize(st));
    }

    // choose the rows-per-block so that a block's local entries fit in shared memory.  A block stores only the entries it
    // sums (one compact record per listed cell, k_pair_masks), i.e. about one double per contribution.
    DevBuf<uint64_t> bk0, bk1, uniq;
    DevBuf<int32_t> flags;
    FDB_TRY(bk0.alloc(nc)); FDB_TRY(bk1.alloc(nc)); FDB_TRY(flags.alloc(nc));
    int rb = compact ? 1024 : 512;
    {
        const double con_per_row = (double)nc / (n > 0 ? n : 1), cells_per_row = (double)s->n_cells / (n > 0 ? n : 1);
        while (rb > 16 && (compact ? 1.1 * rb * con_per_row : 2.2 * rb * cells_per_row * P.ne) * sizeof(double) > smem_target)
            rb = next_rb(rb, fine);
    }
    if (p1tet) {   // finer than the halving ladder: the largest multiple of 8 rows whose local matrices fit the target
        const double per_row = 2.2 * ((double)s->n_cells / (n > 0 ? n : 1)) * P.ne * sizeof(double);
        int fit = (int)(smem_target / per_row) & ~7;
        if (fit > 512) fit = 512;
        if (fit > rb) rb = fit;
    }
    while (rb > 16 && (int64_t)(n + rb - 1) / rb < 8 * s->sm_count) rb = next_rb(rb, fine);  // enough blocks to fill the GPU
    if (const char* e = getenv("FDB_FUSED_RB")) rb = atoi(e) > 0 ? atoi(e) : rb;
    if (p1tet && rb_cap > 0 && rb > rb_cap) rb = rb_cap;
    while (rb_cap > 0 && rb > rb_cap) rb = next_rb(rb, fine);
    for (;; rb = next_rb(rb, fine)) {
        if (rb < 8) return FDB_OK;  // not representable: keep the two-kernel path
        const int nblocks = (n + rb - 1) / rb;
        k_block_cell_keys<<<grid_for(nc, B), B, 0, st>>>(nc, P.ne, shift, rb, ukeys, ids, scan, rank.p, bk0.p);
        FDB_CUDA(cudaGetLastError());
        {
            size_t tb = 0;
            FDB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, bk0.p, bk1.p, (int)nc, 0, 32 + bits_for(nblocks), st));
            DevBuf<char> tmp;
            FDB_TRY(tmp.alloc(tb));
            FDB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, bk0.p, bk1.p, (int)nc, 0, 32 + bits_for(nblocks), st));
            k_flag_heads<<<grid_for(nc, B), B, 0, st>>>(nc, bk1.p, flags.p);
            FDB_CUDA(cudaGetLastError());
            FDB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, flags.p, flags.p, (int)nc, st));
            DevBuf<char> tmp2;
            FDB_TRY(tmp2.alloc(tb));
            FDB_CUDA(cub::DeviceScan::InclusiveSum(tmp2.p, tb, flags.p, flags.p, (int)nc, st));
            FDB_CUDA(cudaStreamSynchronize(st));
        }
        int32_t total = 0;
        FDB_CUDA(cudaMemcpyAsync(&total, flags.p + (nc - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaStreamSynchronize(st));
        FDB_TRY(uniq.alloc(total));
        k_compact_keys<<<grid_for(nc, B), B, 0, st>>>(nc, bk1.p, flags.p, uniq.p);
        FDB_CUDA(cudaGetLastError());
        FDB_TRY(P.f_bcell_ptr.alloc((size_t)nblocks + 1));
        k_block_ptr<<<grid_for(nblocks + 1, B), B, 0, st>>>(nblocks, total, uniq.p, P.f_bcell_ptr.p);
        FDB_CUDA(cudaGetLastError());
        int ecap = 0, lcap = 0;
        int64_t stored = 0;
        DevBuf<int32_t> posof;
        if (!compact) {
            std::vector<int32_t> hptr((size_t)nblocks + 1);
            FDB_CUDA(cudaMemcpyAsync(hptr.data(), P.f_bcell_ptr.p, sizeof(int32_t) * hptr.size(), cudaMemcpyDeviceToHost, st));
            FDB_CUDA(cudaStreamSynchronize(st));
            int max_cells = 0;
            for (int b = 0; b < nblocks; ++b) max_cells = std::max(max_cells, hptr[b + 1] - hptr[b]);
            lcap = (max_cells + 31) / 32 * 32;
            ecap = lcap * P.ne;
            stored = (int64_t)total * P.ne;
            const int64_t smem = (int64_t)ecap * sizeof(double);
            const bool fits = smem <= smem_target || (rb <= 32 && smem <= smem_limit);
            if (!fits || ecap > 65535) continue;
            FDB_TRY(P.f_bcells.alloc(total));
            k_low32<<<grid_for(total, B), B, 0, st>>>(total, uniq.p, P.f_bcells.p);
            FDB_CUDA(cudaGetLastError());
        } else {
            // needed-slot masks, then the listed cells of each block in descending mask order (identical masks adjacent: the
            // threads of a warp then skip the same entries)
            DevBuf<unsigned long long> mask;
            FDB_TRY(mask.alloc(total));
            FDB_CUDA(cudaMemsetAsync(mask.p, 0, sizeof(unsigned long long) * total, st));
            k_pair_masks<<<grid_for(nc, B), B, 0, st>>>(nc, P.ne, shift, rb, ukeys, ids, scan, rank.p, P.f_bcell_ptr.p, uniq.p, mask.p);
            FDB_CUDA(cudaGetLastError());
            DevBuf<uint64_t> mk0, mk1;
            DevBuf<uint32_t> mv0, mv1;
            FDB_TRY(mk0.alloc(total)); FDB_TRY(mk1.alloc(total)); FDB_TRY(mv0.alloc(total)); FDB_TRY(mv1.alloc(total));
            k_mask_keys<<<grid_for(total, B), B, 0, st>>>(total, P.ne, mask.p, mk0.p, mv0.p);
            FDB_CUDA(cudaGetLastError());
            FDB_TRY(radix_sort_pairs(mk0, mk1, mv0, mv1, total, P.ne + 7, st));          // by (count, ~mask) (stable)
            k_block_keys_of<<<grid_for(total, B), B, 0, st>>>(total, mv1.p, uniq.p, mk0.p);
            FDB_CUDA(cudaGetLastError());
            FDB_TRY(radix_sort_pairs(mk0, mk1, mv1, mv0, total, bits_for(nblocks) + 1, st));   // then by block (stable): mv0 = order
            DevBuf<int32_t> len, caps;
            FDB_TRY(len.alloc((size_t)total + 1)); FDB_TRY(posof.alloc(total)); FDB_TRY(caps.alloc(nblocks));
            FDB_TRY(P.f_bcells.alloc(total));
            FDB_TRY(P.f_bmask.alloc((size_t)total + 4));
            k_cell_records<<<grid_for((int64_t)total + 1, B), B, 0, st>>>(total, mv0.p, uniq.p, mask.p, P.f_bcells.p, P.f_bmask.p,
                                                                         len.p, posof.p);
            FDB_CUDA(cudaGetLastError());
            {
                size_t tb = 0;
                FDB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, len.p, len.p, total + 1, st));
                DevBuf<char> tmp;
                FDB_TRY(tmp.alloc(tb));
                FDB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, len.p, len.p, total + 1, st));
            }
            k_block_entry_caps<<<grid_for(nblocks, B), B, 0, st>>>(nblocks, P.f_bcell_ptr.p, len.p, caps.p);
            FDB_CUDA(cudaGetLastError());
            std::vector<int32_t> hcap((size_t)nblocks);
            FDB_CUDA(cudaMemcpyAsync(hcap.data(), caps.p, sizeof(int32_t) * nblocks, cudaMemcpyDeviceToHost, st));
            FDB_CUDA(cudaStreamSynchronize(st));
            int max_ent = 0;
            for (int b = 0; b < nblocks; ++b) { max_ent = std::max(max_ent, hcap[b]); stored += hcap[b]; }
            ecap = (max_ent + 31) / 32 * 32;
            const int64_t smem = (int64_t)ecap * sizeof(double);
            const bool fits = smem <= smem_target || (rb <= 32 && smem <= smem_limit);
            if (!fits || ecap > 65535) continue;
            FDB_TRY(P.f_bbase.alloc((size_t)total + 16));
            k_cell_bases<<<grid_for(total, B), B, 0, st>>>(total, uniq.p, mv0.p, P.f_bcell_ptr.p, len.p, P.f_bbase.p);
            FDB_CUDA(cudaGetLastError());
            FDB_CUDA(cudaStreamSynchronize(st));   // the temporaries of this scope are released below
        }
        const int64_t smem = (int64_t)ecap * sizeof(double);
        FDB_TRY(P.f_lidx.alloc(nc));
        k_gather_index<<<grid_for(nc, B), B, 0, st>>>(nc, P.ne, shift, rb, ukeys, ids, scan, rank.p, P.f_bcell_ptr.p, uniq.p,
                                                     posof.p, P.f_bmask.p, P.f_bbase.p, lcap, P.f_lidx.p);
        FDB_CUDA(cudaGetLastError());
        FDB_CUDA(cudaStreamSynchronize(st));
        P.f_compact = compact;
        P.f_cells_cap = lcap;
        P.f_rb = rb;
        P.f_lcap = ecap;
        P.f_nblocks = nblocks;
        {   // threads per CTA: about two passes of phase 1 over the block's cells, within the limits measured on B200
            const double avg = (double)total / nblocks;
            int nt = 32 * (int)((avg / 2.0 + 31.0) / 32.0);
            const int nt_max = (s->M == 3 && s->R == 1) ? 384 : (p2tet ? 512 : 256);
            P.f_threads = nt < 128 ? 128 : (nt > nt_max ? nt_max : nt);
            if (s->R == 2) P.f_threads = 256;   // many entries per cell: phase 2 dominates the thread count
        }
        P.fused = true;
        if (getenv("FDB_VERBOSE"))
            fprintf(stderr, "[fdb] fused plan: rb=%d blocks=%d entries/block<=%d smem=%lld B cells listed=%d (x%.2f of %d), "
                    "stored entries x%.2f of the contributions\n", rb, nblocks, ecap, (long long)smem, total,
                    (double)total / s->n_cells, s->n_cells, (double)stored / (double)nc);
        return FDB_OK;
    }
}

// ---- fused plan, second stage: everything the kernel reads is laid out block-major, so each CTA streams its own
// contiguous slices (vertex ids, gather indices, segment offsets, destinations) instead of chasing pointers.
// Inside a block the stored entries are ordered by decreasing segment length, so the threads of a warp sum
// segments of (nearly) equal length in phase 2.
// order = 0: by decreasing segment length; 1: by destination (row-major position); 2: length class (log2), then destination
__global__ void k_entry_keys(int64_t nu, int shift, int rb, int order, const uint64_t* __restrict__ ukeys,
                             const int32_t* __restrict__ rank, const int32_t* __restrict__ seg,
                             const int32_t* __restrict__ dst_a, uint64_t* __restrict__ keys, uint32_t* __restrict__ ids) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= nu) return;
    int b = rank[(int)(ukeys[u] >> shift)] / rb;
    int len = seg[u + 1] - seg[u];
    if (len > 65535) len = 65535;
    uint64_t low;
    if (order == 0) low = (uint64_t)(65535 - len) << 31;
    else if (order == 1) low = (uint64_t)dst_a[u];
    else low = ((uint64_t)(31 - (31 - __clz(len))) << 31) | (uint64_t)dst_a[u];
    keys[u] = ((uint64_t)b << 47) | low;   // 16 bits of length / class above 31 bits of destination
    ids[u] = (uint32_t)u;
}

__global__ void k_entry_len(int64_t nu, const uint32_t* __restrict__ sorted_u, const int32_t* __restrict__ seg,
                            int32_t* __restrict__ len) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k > nu) return;
    if (k == nu) { len[k] = 0; return; }
    uint32_t u = sorted_u[k];
    len[k] = seg[u + 1] - seg[u];
}

__global__ void k_block_entry_ptr(int nblocks, int64_t nu, const uint64_t* __restrict__ sorted_keys,
                                  const int32_t* __restrict__ con_off, int32_t* __restrict__ ent_ptr,
                                  int32_t* __restrict__ con_ptr) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nblocks) return;
    uint64_t target = (uint64_t)b << 47;
    int64_t lo = 0, hi = nu;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    ent_ptr[b] = (int32_t)lo;
    con_ptr[b] = con_off[lo];
}

__global__ void k_block_major(int64_t nu, int symmetric, const uint64_t* __restrict__ sorted_keys,
                              const uint32_t* __restrict__ sorted_u, const int32_t* __restrict__ seg,
                              const int32_t* __restrict__ con_off, const int32_t* __restrict__ ent_ptr,
                              const int32_t* __restrict__ con_ptr, const int32_t* __restrict__ dst_a,
                              const int32_t* __restrict__ dst_b, const uint16_t* __restrict__ lidx,
                              int2* __restrict__ dst_bm, int32_t* __restrict__ dst1_bm, uint16_t* __restrict__ segrel,
                              uint16_t* __restrict__ lidx_bm) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= nu) return;
    const int b = (int)(sorted_keys[k] >> 47);
    const uint32_t u = sorted_u[k];
    const int t0 = seg[u], t1 = seg[u + 1];
    const int c = con_off[k];
    if (symmetric) dst_bm[k] = make_int2(dst_a[u], dst_b[u]);
    else dst1_bm[k] = dst_a[u];
    segrel[k + b] = (uint16_t)(c - con_ptr[b]);
    if ((int)k + 1 == ent_ptr[b + 1]) segrel[k + b + 1] = (uint16_t)(con_ptr[b + 1] - con_ptr[b]);
    for (int t = t0; t < t1; ++t) lidx_bm[c + (t - t0)] = lidx[t];
}

__global__ void k_block_verts(int64_t total, int nv, int n_cells, const int32_t* __restrict__ bcells,
                              const int32_t* __restrict__ verts, int32_t* __restrict__ bverts) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    int e = bcells[i];
    for (int k = 0; k < nv; ++k) bverts[i * nv + k] = verts[(size_t)k * n_cells + e];
}

// ---- bank-aware numbering of a block's cells (P1 elements: default on tetrahedra, FDB_FUSED_BANKS=0/1) ------------
// Phase 2 of the fused kernel reads loc[slot * lcap + cell] with 16 lanes per pass; lcap is a multiple of 16, so the
// bank pair of a read is (cell mod 16) and sixteen effectively random cells collide like balls into bins: 5.2
// wavefronts per 64-bit load (ncu; reproduced by tools/plan_model.py).  Here the cells keep their 16-cell window of the
// list -- so the lanes of phase 1 still gather neighbouring cells and store conflict-free -- but inside each window the
// position (= bank) of every cell is chosen greedily against the cells it is read together with: for each window in
// order, for each cell in order, take the free bank with the fewest cells already placed there among its co-readers
// ((half-warp, step) groups of phase 2).  The model gives 3.9 wavefronts per load.  Only where a value is parked in
// shared memory changes; every entry still adds the same values in the same order.
__global__ void k_max_seg_len(int64_t nu, const int32_t* __restrict__ seg, int* __restrict__ out) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= nu) return;
    atomicMax(out, seg[u + 1] - seg[u]);
}

__global__ void __launch_bounds__(128)
k_bank_colour(int lcap, int lmax, int max_con, const int32_t* __restrict__ meta, uint16_t* __restrict__ lidx,
              const uint16_t* __restrict__ segrel, uint16_t* __restrict__ newpos_g) {
    extern __shared__ unsigned char dyn[];
    const int b = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const int32_t* m = meta + (size_t)b * 8;
    const int c0 = m[0], ncon = m[1], e0 = m[2], ne_b = m[3], cc0 = m[4], ncell = m[5];
    const int n_half = (ne_b + 15) >> 4;
    const int G = n_half * lmax;
    // shared layout: cnt[lcap] | off[lcap + 1] | newpos[lcap] (int32) | items[max_con] (uint16) | hist[G * 16] (uint8)
    int* cnt = reinterpret_cast<int*>(dyn);
    int* off = cnt + lcap;
    int* newpos = off + lcap + 1;
    uint16_t* items = reinterpret_cast<uint16_t*>(newpos + lcap);
    unsigned char* hist = reinterpret_cast<unsigned char*>(items + ((max_con + 1) & ~1));
    for (int i = tid; i < lcap; i += NT) { cnt[i] = 0; newpos[i] = i; }
    for (int i = tid; i < G * 16; i += NT) hist[i] = 0;
    __syncthreads();
    const uint16_t* sr = segrel + e0 + b;
    for (int k = tid; k < ne_b; k += NT)
        for (int t = sr[k]; t < sr[k + 1]; ++t) atomicAdd(&cnt[lidx[c0 + t] % lcap], 1);
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < lcap; ++i) { off[i] = acc; acc += cnt[i]; cnt[i] = 0; }
        off[lcap] = acc;
    }
    __syncthreads();
    for (int k = tid; k < ne_b; k += NT) {
        const int t0 = sr[k];
        for (int t = t0; t < sr[k + 1]; ++t) {
            const int lc = lidx[c0 + t] % lcap;
            items[off[lc] + atomicAdd(&cnt[lc], 1)] = (uint16_t)((k >> 4) * lmax + (t - t0));
        }
    }
    __syncthreads();
    if (tid < 32) {  // one warp places the cells, lane k < 16 prices bank k
        const int k = tid & 15;
        for (int w0 = 0; w0 < ncell; w0 += 16) {
            const int mwin = min(16, ncell - w0);
            unsigned freemask = (mwin == 16) ? 0xffffu : ((1u << mwin) - 1u);
            for (int i = 0; i < mwin; ++i) {
                const int c = w0 + i;
                int cost = 1 << 20;
                if ((freemask >> k) & 1u) {
                    cost = 0;
                    for (int x = off[c]; x < off[c + 1]; ++x) cost += hist[(int)items[x] * 16 + k];
                }
                int best = (cost << 5) | k;  // ties go to the lowest bank
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
                const int kb = best & 31;
                freemask &= ~(1u << kb);
                if (tid == 0) {
                    newpos[c] = w0 + kb;
                    for (int x = off[c]; x < off[c + 1]; ++x) hist[(int)items[x] * 16 + kb] += 1;
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();
    for (int t = tid; t < ncon; t += NT) {
        const int v = lidx[c0 + t];
        lidx[c0 + t] = (uint16_t)((v / lcap) * lcap + newpos[v % lcap]);
    }
    for (int i = tid; i < ncell; i += NT) newpos_g[cc0 + i] = (uint16_t)newpos[i];
}

__global__ void k_apply_cell_perm(int nv, const int32_t* __restrict__ meta, const uint16_t* __restrict__ newpos_g,
                                  const int32_t* __restrict__ bverts, const int32_t* __restrict__ bcells,
                                  int32_t* __restrict__ bverts2, int32_t* __restrict__ bcells2) {
    const int32_t* m = meta + (size_t)blockIdx.x * 8;
    const int cc0 = m[4], ncell = m[5];
    for (int i = threadIdx.x; i < ncell; i += blockDim.x) {
        const int j = newpos_g[cc0 + i];
        for (int v = 0; v < nv; ++v) bverts2[(size_t)(cc0 + j) * nv + v] = bverts[(size_t)(cc0 + i) * nv + v];
        bcells2[cc0 + j] = bcells[cc0 + i];
    }
}

static int bank_colour_plan(fdb_space* s, Pattern& P) {
    cudaStream_t st = s->stream;
    DevBuf<int> dmax;
    FDB_TRY(dmax.alloc(1));
    FDB_CUDA(cudaMemsetAsync(dmax.p, 0, sizeof(int), st));
    k_max_seg_len<<<grid_for(P.n_unique, 256), 256, 0, st>>>(P.n_unique, P.seg.p, dmax.p);
    FDB_CUDA(cudaGetLastError());
    int lmax = 0;
    FDB_CUDA(cudaMemcpyAsync(&lmax, dmax.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    const int n_half = (P.f_max_ent + 15) / 16;
    const int64_t G = (int64_t)n_half * lmax;
    const int lcap = P.f_cells_cap;   // cells per slot row of the block's local-matrix array
    const size_t smem = sizeof(int) * ((size_t)3 * lcap + 1) + sizeof(uint16_t) * (((size_t)P.f_max_con + 1) & ~(size_t)1) +
                        (size_t)G * 16 + 16;
    if (G <= 0 || G > 65535 || smem > 200 * 1024) return FDB_OK;  // lists too long for the shared-memory tables: keep the order
    FDB_CUDA(cudaFuncSetAttribute(k_bank_colour, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t total = (int64_t)P.f_bcells.n;
    const int nv = s->M + 1;
    DevBuf<uint16_t> newpos;
    DevBuf<int32_t> bverts2, bcells2;
    FDB_TRY(newpos.alloc((size_t)total));
    FDB_TRY(bverts2.alloc((size_t)total * nv));
    FDB_TRY(bcells2.alloc((size_t)total));
    k_bank_colour<<<P.f_nblocks, 128, smem, st>>>(lcap, lmax, P.f_max_con, P.f_meta.p, P.f_lidx.p, P.f_segrel.p, newpos.p);
    FDB_CUDA(cudaGetLastError());
    k_apply_cell_perm<<<P.f_nblocks, 128, 0, st>>>(nv, P.f_meta.p, newpos.p, P.f_bverts.p, P.f_bcells.p, bverts2.p, bcells2.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(st));
    std::swap(P.f_bverts.p, bverts2.p);
    std::swap(P.f_bverts.n, bverts2.n);
    std::swap(P.f_bcells.p, bcells2.p);
    std::swap(P.f_bcells.n, bcells2.n);
    return FDB_OK;
}


// block-local node copies: (block, node) key of every vertex of every listed cell (one CTA per block)
__global__ void k_block_node_keys(const int32_t* __restrict__ bcell_ptr, int nv, const int32_t* __restrict__ bverts,
                                  uint64_t* __restrict__ keys) {
    const int b = blockIdx.x;
    const int64_t t0 = (int64_t)bcell_ptr[b] * nv, t1 = (int64_t)bcell_ptr[b + 1] * nv;
    for (int64_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) keys[t] = ((uint64_t)b << 32) | (uint32_t)bverts[t];
}
// index of every vertex in its block's ascending node list (4 per listed cell; unused slots 0)
__global__ void k_block_local_ids(const int32_t* __restrict__ bcell_ptr, int nv, const int32_t* __restrict__ bverts,
                                  const uint64_t* __restrict__ uniq, const int32_t* __restrict__ node_ptr,
                                  uint16_t* __restrict__ bvloc) {
    const int b = blockIdx.x;
    const int n0 = node_ptr[b], n1 = node_ptr[b + 1];
    for (int c = bcell_ptr[b] + threadIdx.x; c < bcell_ptr[b + 1]; c += blockDim.x) {
        for (int k = 0; k < 4; ++k) {
            int id = 0;
            if (k < nv) {
                const uint64_t key = ((uint64_t)b << 32) | (uint32_t)bverts[(int64_t)c * nv + k];
                int lo = n0, hi = n1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (uniq[mid] < key) lo = mid + 1;
                    else hi = mid;
                }
                id = lo - n0;
            }
            bvloc[(int64_t)c * 4 + k] = (uint16_t)id;
        }
    }
}
__global__ void k_block_coords(int64_t n, int pk, const uint64_t* __restrict__ uniq, const double* __restrict__ coords_pk,
                               double* __restrict__ bxy, double* __restrict__ bz) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* c = coords_pk + (size_t)(uniq[i] & 0xffffffffu) * pk;
    bxy[2 * i] = c[0];
    bxy[2 * i + 1] = c[1];
    if (pk == 4) bz[i] = c[2];
}

// fills P.f_bvloc / P.f_bcoords and the per-block node ranges (first node, count) of hnode
static int build_block_nodes(fdb_space* s, Pattern& P, std::vector<int32_t>& hnode) {
    cudaStream_t st = s->stream;
    const int B = 256, nblocks = P.f_nblocks, nv = s->M + 1, pk = (s->N == 3) ? 4 : 2;
    const int64_t total = (int64_t)P.f_bcells.n, tv = total * nv;
    DevBuf<uint64_t> k0, k1, uniq;
    DevBuf<int32_t> flags, node_ptr;
    FDB_TRY(k0.alloc(tv)); FDB_TRY(k1.alloc(tv)); FDB_TRY(flags.alloc(tv)); FDB_TRY(node_ptr.alloc((size_t)nblocks + 1));
    k_block_node_keys<<<nblocks, B, 0, st>>>(P.f_bcell_ptr.p, nv, P.f_bverts.p, k0.p);
    FDB_CUDA(cudaGetLastError());
    {
        size_t tb = 0;
        FDB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, k0.p, k1.p, (int)tv, 0, 32 + bits_for(nblocks), st));
        DevBuf<char> tmp;
        FDB_TRY(tmp.alloc(tb));
        FDB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, k0.p, k1.p, (int)tv, 0, 32 + bits_for(nblocks), st));
        k_flag_heads<<<grid_for(tv, B), B, 0, st>>>(tv, k1.p, flags.p);
        FDB_CUDA(cudaGetLastError());
        FDB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, flags.p, flags.p, (int)tv, st));
        DevBuf<char> tmp2;
        FDB_TRY(tmp2.alloc(tb));
        FDB_CUDA(cub::DeviceScan::InclusiveSum(tmp2.p, tb, flags.p, flags.p, (int)tv, st));
        FDB_CUDA(cudaStreamSynchronize(st));
    }
    int32_t n_nodes_total = 0;
    FDB_CUDA(cudaMemcpyAsync(&n_nodes_total, flags.p + (tv - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    FDB_TRY(uniq.alloc(n_nodes_total));
    k_compact_keys<<<grid_for(tv, B), B, 0, st>>>(tv, k1.p, flags.p, uniq.p);
    FDB_CUDA(cudaGetLastError());
    k_block_ptr<<<grid_for(nblocks + 1, B), B, 0, st>>>(nblocks, n_nodes_total, uniq.p, node_ptr.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(P.f_bvloc.alloc((size_t)total * 4 + 8));
    k_block_local_ids<<<nblocks, B, 0, st>>>(P.f_bcell_ptr.p, nv, P.f_bverts.p, uniq.p, node_ptr.p, P.f_bvloc.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(P.f_bcoords.alloc((size_t)n_nodes_total * 2 + 8));
    FDB_TRY(P.f_bz.alloc((size_t)(pk == 4 ? n_nodes_total : 0) + 8));
    k_block_coords<<<grid_for(n_nodes_total, B), B, 0, st>>>(n_nodes_total, pk, uniq.p, s->coords_pk.p, P.f_bcoords.p, P.f_bz.p);
    FDB_CUDA(cudaGetLastError());
    hnode.resize((size_t)nblocks + 1);
    FDB_CUDA(cudaMemcpyAsync(hnode.data(), node_ptr.p, sizeof(int32_t) * hnode.size(), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    int max_nodes = 0;
    for (int b = 0; b < nblocks; ++b) max_nodes = std::max(max_nodes, hnode[b + 1] - hnode[b]);
    if (max_nodes > 65535) {   // 16-bit block-local indices
        P.f_nodes = false;
        return FDB_OK;
    }
    P.f_node_z_off = max_nodes * 16;                                             // (x, y) pairs
    P.f_node_bytes = P.f_node_z_off + (pk == 4 ? ((max_nodes + 2) * 8 + 15) / 16 * 16 : 0);   // + z (copied from an even node index)
    if (getenv("FDB_VERBOSE"))
        fprintf(stderr, "[fdb] fused plan: block-local nodes: %d in all (x%.2f of %d), <= %d per block (%d B of shared memory)\n",
                n_nodes_total, (double)n_nodes_total / s->n_nodes, s->n_nodes, max_nodes, P.f_node_bytes);
    return FDB_OK;
}

static int finish_fused_plan(fdb_space* s, Pattern& P, int shift, const uint64_t* ukeys, const int32_t* rank) {
    cudaStream_t st = s->stream;
    const int B = 256, rb = P.f_rb, nblocks = P.f_nblocks;
    const int64_t nu = P.n_unique;
    DevBuf<uint64_t> ek0, ek1;
    DevBuf<uint32_t> eu0, eu1;
    DevBuf<int32_t> con_off;
    FDB_TRY(ek0.alloc(nu)); FDB_TRY(ek1.alloc(nu)); FDB_TRY(eu0.alloc(nu)); FDB_TRY(eu1.alloc(nu));
    FDB_TRY(con_off.alloc((size_t)nu + 1));
    // entries of a block: by decreasing segment length (threads of a warp sum segments of equal length); P2 tetrahedra: by
    // length class, then destination (measured on B200: -1.5 % there, +2.5 % on P1 tetrahedra, neutral on triangles)
    int order = (s->M == 3 && s->R == 2) ? 2 : 0;
    if (const char* e = getenv("FDB_FUSED_ENTRY_ORDER")) order = atoi(e);
    k_entry_keys<<<grid_for(nu, B), B, 0, st>>>(nu, shift, rb, order, ukeys, rank, P.seg.p, P.dst_a.p, ek0.p, eu0.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(radix_sort_pairs(ek0, ek1, eu0, eu1, nu, 47 + bits_for(nblocks), st));
    k_entry_len<<<grid_for(nu + 1, B), B, 0, st>>>(nu, eu1.p, P.seg.p, con_off.p);
    FDB_CUDA(cudaGetLastError());
    {
        size_t tb = 0;
        FDB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, con_off.p, con_off.p, (int)nu + 1, st));
        DevBuf<char> tmp;
        FDB_TRY(tmp.alloc(tb));
        FDB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, con_off.p, con_off.p, (int)nu + 1, st));
        FDB_CUDA(cudaStreamSynchronize(st));
    }
    FDB_TRY(P.f_ent_ptr.alloc((size_t)nblocks + 1));
    FDB_TRY(P.f_con_ptr.alloc((size_t)nblocks + 1));
    k_block_entry_ptr<<<grid_for(nblocks + 1, B), B, 0, st>>>(nblocks, nu, ek1.p, con_off.p, P.f_ent_ptr.p,
                                                             P.f_con_ptr.p);
    FDB_CUDA(cudaGetLastError());
    std::vector<int32_t> he((size_t)nblocks + 1), hc((size_t)nblocks + 1);
    FDB_CUDA(cudaMemcpyAsync(he.data(), P.f_ent_ptr.p, sizeof(int32_t) * he.size(), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(hc.data(), P.f_con_ptr.p, sizeof(int32_t) * hc.size(), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    int max_ent = 0, max_con = 0;
    for (int b = 0; b < nblocks; ++b) {
        max_ent = std::max(max_ent, he[b + 1] - he[b]);
        max_con = std::max(max_con, hc[b + 1] - hc[b]);
    }
    if (max_con > 65535) {  // segment offsets are 16-bit
        P.fused = false;
        return FDB_OK;
    }
    P.f_max_ent = max_ent;
    P.f_max_con = max_con;
    if (P.symmetric) FDB_TRY(P.f_dst.alloc((size_t)P.n_unique + 4));    // + slack for the 16-byte bulk-copy granules
    else FDB_TRY(P.f_dst1.alloc((size_t)P.n_unique + 8));
    FDB_TRY(P.f_segrel.alloc((size_t)P.n_unique + nblocks + 1 + 16));
    DevBuf<uint16_t> lidx_bm;
    FDB_TRY(lidx_bm.alloc((size_t)P.n_contrib + 16));
    k_block_major<<<grid_for(nu, B), B, 0, st>>>(nu, P.symmetric ? 1 : 0, ek1.p, eu1.p, P.seg.p, con_off.p,
                                                P.f_ent_ptr.p, P.f_con_ptr.p, P.dst_a.p, P.dst_b.p, P.f_lidx.p,
                                                P.f_dst.p, P.f_dst1.p, P.f_segrel.p, lidx_bm.p);
    FDB_CUDA(cudaGetLastError());
    const int64_t total = (int64_t)P.f_bcells.n;
    // vertex ids of the listed cells, block-major (streamed by the CTA instead of gathered through the cell id)
    const int nv = s->M + 1;
    FDB_TRY(P.f_bverts.alloc((size_t)total * nv));
    k_block_verts<<<grid_for(total, B), B, 0, st>>>(total, nv, s->n_cells, P.f_bcells.p, s->verts_p, P.f_bverts.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(st));
    // per-block descriptor: {first contribution, contributions, first entry, entries, first listed cell, cells, first node, nodes}
    {
        std::vector<int32_t> hcell((size_t)nblocks + 1);
        FDB_CUDA(cudaMemcpyAsync(hcell.data(), P.f_bcell_ptr.p, sizeof(int32_t) * hcell.size(), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaStreamSynchronize(st));
        P.f_max_cells = 0;
        std::vector<int32_t> meta((size_t)nblocks * 8, 0);
        for (int b = 0; b < nblocks; ++b) {
            int32_t* m = meta.data() + (size_t)b * 8;
            m[0] = hc[b]; m[1] = hc[b + 1] - hc[b]; m[2] = he[b]; m[3] = he[b + 1] - he[b];
            m[4] = hcell[b]; m[5] = hcell[b + 1] - hcell[b];
            P.f_max_cells = std::max(P.f_max_cells, m[5]);
        }
        FDB_TRY(P.f_meta.alloc(meta.size()));
        FDB_CUDA(cudaMemcpyAsync(P.f_meta.p, meta.data(), sizeof(int32_t) * meta.size(), cudaMemcpyHostToDevice, st));
        FDB_CUDA(cudaStreamSynchronize(st));
        // swap in the block-major gather indices (DevBuf is not copyable: exchange the raw pointers)
        std::swap(P.f_lidx.p, lidx_bm.p);
        std::swap(P.f_lidx.n, lidx_bm.n);
        // bank-aware positions of the cells inside their 16-cell windows (slot-major layout only; permutes the listed cells)
        // default on P1 tetrahedra (C4 0.301 -> 0.295 ms under the persistent kernel, whose limiter is the shared-memory
        // wavefront pipe; no effect on P1 triangles); FDB_FUSED_BANKS = 0 / 1 forces it
        {
            const char* e = getenv("FDB_FUSED_BANKS");
            if (!P.f_compact && (e ? atoi(e) != 0 : (s->M == 3 && s->R == 1))) FDB_TRY(bank_colour_plan(s, P));
        }
        // block-local node copies (P1 elements; FDB_FUSED_NODES=0 switches them off): measured -6 % on C2 with the plain fused
        // kernel, and the persistent kernel of P1 tetrahedra is built on them (C4 0.414 -> 0.328 ms)
        std::vector<int32_t> hnode;
        P.f_nodes = s->M == s->N && !(getenv("FDB_FUSED_NODES") != nullptr && atoi(getenv("FDB_FUSED_NODES")) == 0);
        // P2: the node copies only serve the persistent kernel -- the default on triangles (C3 0.275 -> 0.240 ms), measured
        // slower on tetrahedra (0.79 -> 0.87 ms: the extra lists cost a resident CTA); FDB_FUSED_PERSIST_P2 = 0 / 1 forces it
        if (s->R == 2) {
            const char* e = getenv("FDB_FUSED_PERSIST_P2");
            if (!(e ? atoi(e) != 0 : s->M == 2)) P.f_nodes = false;
        }
        if (P.f_nodes) FDB_TRY(build_block_nodes(s, P, hnode));
        if (P.f_nodes) {
            for (int b = 0; b < nblocks; ++b) {
                meta[(size_t)b * 8 + 6] = hnode[b];
                meta[(size_t)b * 8 + 7] = hnode[b + 1] - hnode[b];
            }
            FDB_CUDA(cudaMemcpyAsync(P.f_meta.p, meta.data(), sizeof(int32_t) * meta.size(), cudaMemcpyHostToDevice, st));
            FDB_CUDA(cudaStreamSynchronize(st));
        }
    }
    return FDB_OK;
}

// sorted contribution ids (inverse of the scatter map) and the entry id of every contribution, rebuilt on demand
__global__ void k_invert_pos(int64_t nc, int n_cells, int ne, const int32_t* __restrict__ pos, uint32_t* __restrict__ ids) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // k = s * n_cells + e
    if (k >= nc) return;
    int s = (int)(k / n_cells), e = (int)(k % n_cells);
    ids[pos[k]] = (uint32_t)((int64_t)e * ne + s);
}
__global__ void k_fill_uid(int64_t nu, const int32_t* __restrict__ seg, int32_t* __restrict__ scan) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= nu) return;
    for (int t = seg[u]; t < seg[u + 1]; ++t) scan[t] = (int32_t)u + 1;  // same convention as the inclusive scan
}

// The fused plan costs about as much as the pattern itself, so it is built only when a pattern is assembled a second
// time (stiffness + mass, time stepping, repeated solves) or when the caller asks for it (fdb_space_prepare).
int ensure_fused_plan(fdb_space* s, Pattern* Pp) {
    Pattern& P = *Pp;
    if (P.fused_tried) return FDB_OK;
    P.fused_tried = true;
    cudaStream_t st = s->stream;
    const int B = 256;
    DevBuf<uint32_t> ids;
    DevBuf<int32_t> scan, rank;
    FDB_TRY(ids.alloc(P.n_contrib));
    FDB_TRY(scan.alloc(P.n_contrib));
    k_invert_pos<<<grid_for(P.n_contrib, B), B, 0, st>>>(P.n_contrib, s->n_cells, P.ne, P.pos.p, ids.p);
    k_fill_uid<<<grid_for(P.n_unique, B), B, 0, st>>>(P.n_unique, P.seg.p, scan.p);
    FDB_CUDA(cudaGetLastError());
    // static shared memory of the kernel: mbarrier + the staged reference-tensor rows (none in the lean P1 stiffness form)
    const size_t stat = 64 + ((s->R == 1 && s->M == 3) ? 0 : sizeof(double) * s->nb * s->nb * tens_stride(s->M));
    const size_t smem_sm = 228 * 1024, smem_cta_max = 227 * 1024, reserve = 1024 + stat;  // per-CTA reservation + static
    for (int rb_cap = 0;;) {
        FDB_TRY(build_fused_plan(s, P, P.shift, P.ukeys.p, ids.p, scan.p, rank, rb_cap));
        if (!P.fused) break;
        FDB_TRY(finish_fused_plan(s, P, P.shift, P.ukeys.p, rank.p));
        if (!P.fused) {   // more than 65535 contributions in one block (16-bit segment offsets): smaller blocks
            rb_cap = next_rb(P.f_rb, true);
            if (rb_cap < 8) break;
            continue;
        }
        if (s->M == 3 && s->R == 1 && P.f_nodes && !persist_fits(P) && P.f_rb > 24 && !getenv("FDB_FUSED_RB")) {
            rb_cap = (P.f_rb - 8) & ~7;   // P1 tetrahedra: the persistent kernel needs two CTAs per SM -- a few rows less
            continue;
        }
        const size_t plain = fused_smem_bytes(P, false), with_dst = fused_smem_bytes(P, true);
        if (plain + reserve > smem_cta_max) {   // the block lists do not fit beside the local matrices: smaller blocks
            P.fused = false;
            rb_cap = next_rb(P.f_rb, true);
            if (rb_cap < 8) break;
            continue;
        }
        // destinations in shared memory when that costs no resident CTA
        auto ctas = [&](size_t dyn) {
            const size_t by_smem = smem_sm / (dyn + reserve), by_threads = (size_t)(2048 / P.f_threads);
            return by_smem < by_threads ? by_smem : by_threads;
        };
        P.f_dsm = with_dst + reserve <= smem_cta_max && ctas(with_dst) == ctas(plain);
        if (const char* e = getenv("FDB_FUSED_DSM")) P.f_dsm = atoi(e) != 0 && with_dst + reserve <= smem_cta_max;
        if (getenv("FDB_VERBOSE"))
            fprintf(stderr, "[fdb] fused plan: threads=%d smem=%zu B (%zu with destinations), dst in smem=%d, max entries=%d "
                    "contributions=%d\n", P.f_threads, plain, with_dst, (int)P.f_dsm, P.f_max_ent, P.f_max_con);
        break;
    }
    FDB_CUDA(cudaStreamSynchronize(st));
    return FDB_OK;
}

// ---- row-wise pattern build ---------------------------------------------------------------------------------------
// Same arrays as the sort-based build below, bit for bit, without sorting the emitted triplets (101 M 64-bit keys on C4).
// (1) The (cell, local index) incidences of every dof, by one stable 32-bit sort: cells ascend inside a dof.
// (2) One warp per row r: the distinct dofs of the cells incident to r are collected and ranked in shared memory -- that
//     is row r of the full pattern, and its prefix (columns <= r) the stored entries of a symmetric operator
//     (fem_assembler.h:96).  Row sizes are scanned into rowptr / entry / contribution offsets.
// (3) One warp per row again (the row's columns come back from a scratch list): every emitted triplet of the row finds
//     its column by binary search; a counting pass sizes
//     the segments, and a second pass over the triplets in emission order (cells ascending, as setFromTriplets sees them,
//     fem_assembler.h:112) gives each its place inside its segment -- a stable counting sort, so the left-to-right sum
//     of a segment is still Eigen's duplicate order.
// (4) Symmetric operators: the mirror position of entry (r, c) is found in row c of the full pattern.
// Rows with more than RW_UCAP distinct columns (never seen on a simplicial mesh) send the whole build to the sort path.
constexpr int RW_UCAP = 256, RW_WARPS = 8;
struct RowArgs {
    int nb, n_cells, symmetric;
    const int32_t* dofs;      // SoA [nb][n_cells]
    const uint32_t* inc;      // cell * nb + local index, sorted by (dof, cell)
    const int32_t* inc_ptr;   // n_dofs + 1
};

__device__ __forceinline__ uint32_t rw_candidate(const RowArgs& a, int i0, int k, int& cell, int& ai, int& j) {
    const int q = k / a.nb;
    j = k - q * a.nb;
    const uint32_t cid = __ldg(a.inc + i0 + q);
    cell = (int)(cid / (uint32_t)a.nb);
    ai = (int)(cid - (uint32_t)cell * (uint32_t)a.nb);
    return (uint32_t)__ldg(a.dofs + (size_t)j * a.n_cells + cell);
}

// distinct columns of row r, ascending, in S[0, u); returns u, or -1 when they do not fit.  *low_cand counts (per lane)
// the emitted triplets with column <= r.
__device__ int rw_row_columns(const RowArgs& a, int r, int i0, int total, int lane, uint32_t* U, uint32_t* S, int* low_cand) {
    int u = 0, lc = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int k0 = 0; k0 < total; k0 += 32) {
        const int k = k0 + lane;
        const bool valid = k < total;
        uint32_t col = 0;
        if (valid) {
            int c_, a_, j_;
            col = rw_candidate(a, i0, k, c_, a_, j_);
            lc += col <= (uint32_t)r;
        }
        bool fresh = valid;
        for (int q = 0; q < u; ++q) fresh = fresh && (U[q] != col);
        const unsigned newm = __ballot_sync(0xffffffffu, fresh);
        bool lead = false;
        if (fresh) {
            const unsigned m = __match_any_sync(newm, col);
            lead = (m & lt) == 0;
        }
        const unsigned leadm = __ballot_sync(0xffffffffu, lead);
        const int add = __popc(leadm);
        if (u + add > RW_UCAP) return -1;
        if (lead) U[u + __popc(leadm & lt)] = col;
        u += add;
        __syncwarp();
    }
    for (int e = lane; e < u; e += 32) {   // the columns are distinct: rank = number of smaller ones
        const uint32_t v = U[e];
        int rk = 0;
        for (int q = 0; q < u; ++q) rk += U[q] < v;
        S[rk] = v;
    }
    __syncwarp();
    *low_cand = lc;
    return u;
}

__device__ __forceinline__ int rw_warp_sum(int v) {
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// cnt3 = [full row sizes | stored entries per row | emitted triplets per row], each n + 1 long (last element stays 0)
__global__ void __launch_bounds__(32 * RW_WARPS)
k_row_counts(int n, RowArgs a, int32_t* __restrict__ full_cnt, int32_t* __restrict__ low_cnt, int32_t* __restrict__ con_cnt,
             uint32_t* __restrict__ scratch, int* __restrict__ overflow) {
    __shared__ uint32_t sU[RW_WARPS][RW_UCAP], sS[RW_WARPS][RW_UCAP];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * RW_WARPS + w;
    if (r >= n) return;
    const int i0 = a.inc_ptr[r], total = (a.inc_ptr[r + 1] - i0) * a.nb;
    int lc = 0;
    const int u = rw_row_columns(a, r, i0, total, lane, sU[w], sS[w], &lc);
    if (u < 0) {
        if (lane == 0) *overflow = 1;
        return;
    }
    int low = 0;
    uint32_t* keep = scratch + (size_t)i0 * a.nb;   // the row's sorted columns, kept for k_row_fill (u <= total)
    for (int e = lane; e < u; e += 32) {
        const uint32_t v = sS[w][e];
        keep[e] = v;
        low += v <= (uint32_t)r;
    }
    low = rw_warp_sum(low);
    lc = rw_warp_sum(lc);
    if (lane == 0) {
        full_cnt[r] = u;
        low_cnt[r] = a.symmetric ? low : u;
        con_cnt[r] = a.symmetric ? lc : total;
    }
}

__device__ __forceinline__ int rw_find(const uint32_t* S, int n, uint32_t v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (S[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

constexpr int RW_K = 4;   // rows with at most 32 * RW_K emitted triplets keep them in registers between the passes

__device__ __forceinline__ int rw_slot(const RowArgs& a, int ai, int j) {
    if (a.symmetric) {
        const int i = ai < j ? ai : j, jj = ai < j ? j : ai;
        return i * a.nb - i * (i - 1) / 2 + (jj - i);
    }
    return ai * a.nb + j;
}

// exclusive scan of the per-column counts in run[0, nl) (in place); the segment starts go to seg
__device__ __forceinline__ void rw_scan_segments(int* run, int nl, int lane, int c0, int32_t* __restrict__ seg_row) {
    int carry = 0;
    for (int k0 = 0; k0 < nl; k0 += 32) {
        const int k = k0 + lane;
        const int c = k < nl ? run[k] : 0;
        int inc = c;
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (k < nl) {
            run[k] = carry + inc - c;
            seg_row[k] = c0 + carry + inc - c;
        }
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// place of one batch of triplets inside their segments: lanes ascend with the emission order, so the rank inside a group
// of equal columns (match_any) plus the running count of the column is the stable position
__device__ __forceinline__ int rw_place(int* run, bool act, int kk, int lane) {
    const unsigned lt = (1u << lane) - 1u;
    const unsigned actm = __ballot_sync(0xffffffffu, act);
    unsigned m = 0;
    int base = 0;
    if (act) {
        m = __match_any_sync(actm, kk);
        base = run[kk];
    }
    __syncwarp();
    if (act && (m >> lane) == 1u) run[kk] = base + __popc(m);   // highest lane of the group
    __syncwarp();
    return base + __popc(m & lt);
}

__global__ void __launch_bounds__(32 * RW_WARPS)
k_row_fill(int n, RowArgs a, int shift, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ lptr,
           const int32_t* __restrict__ cptr, const uint32_t* __restrict__ scratch, int32_t* __restrict__ colidx,
           uint64_t* __restrict__ ukeys, int32_t* __restrict__ dst_a, int32_t* __restrict__ seg, int32_t* __restrict__ pos) {
    __shared__ uint32_t sS[RW_WARPS][RW_UCAP];
    __shared__ int sRun[RW_WARPS][RW_UCAP];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * RW_WARPS + w;
    if (r >= n) return;
    const int i0 = a.inc_ptr[r], total = (a.inc_ptr[r + 1] - i0) * a.nb;
    const int p0 = rowptr[r], u = rowptr[r + 1] - p0, l0 = lptr[r], nl = lptr[r + 1] - l0, c0 = cptr[r];
    uint32_t* S = sS[w];
    int* run = sRun[w];
    const uint32_t* kept = scratch + (size_t)i0 * a.nb;   // sorted distinct columns of the row (k_row_counts)
    for (int k = lane; k < u; k += 32) {
        const uint32_t c = kept[k];
        S[k] = c;
        colidx[p0 + k] = (int32_t)c;
        if (k < nl) {
            ukeys[l0 + k] = ((uint64_t)(uint32_t)r << shift) | c;
            dst_a[l0 + k] = p0 + k;     // the stored columns (<= r) are the head of the full row
            run[k] = 0;
        }
    }
    if (r == n - 1 && lane == 0) seg[lptr[n]] = cptr[n];
    __syncwarp();
    if (total <= 32 * RW_K) {
        // every triplet of the row is read once: column rank and scatter-map index stay in registers
        int kk[RW_K];
        uint32_t pidx[RW_K];
#pragma unroll
        for (int b = 0; b < RW_K; ++b) {
            const int k = 32 * b + lane;
            kk[b] = -1;
            pidx[b] = 0;
            if (k < total) {
                int cell, ai, j;
                const uint32_t col = rw_candidate(a, i0, k, cell, ai, j);
                if (!a.symmetric || col <= (uint32_t)r) {
                    kk[b] = rw_find(S, nl, col);
                    pidx[b] = (uint32_t)rw_slot(a, ai, j) * (uint32_t)a.n_cells + (uint32_t)cell;
                    atomicAdd(&run[kk[b]], 1);
                }
            }
        }
        __syncwarp();
        rw_scan_segments(run, nl, lane, c0, seg + l0);
        __syncwarp();
#pragma unroll
        for (int b = 0; b < RW_K; ++b) {
            if (32 * b < total) {   // warp-uniform
                const bool act = kk[b] >= 0;
                const int at = rw_place(run, act, act ? kk[b] : 0, lane);
                if (act) pos[pidx[b]] = c0 + at;
            }
        }
        return;
    }
    // general rows: three passes over the triplets (segment sizes, scan, places)
    for (int k0 = 0; k0 < total; k0 += 32) {
        const int k = k0 + lane;
        if (k < total) {
            int cell, ai, j;
            const uint32_t col = rw_candidate(a, i0, k, cell, ai, j);
            if (!a.symmetric || col <= (uint32_t)r) atomicAdd(&run[rw_find(S, nl, col)], 1);
        }
    }
    __syncwarp();
    rw_scan_segments(run, nl, lane, c0, seg + l0);
    __syncwarp();
    for (int k0 = 0; k0 < total; k0 += 32) {
        const int k = k0 + lane;
        int cell = 0, ai = 0, j = 0, kk = 0;
        bool act = false;
        if (k < total) {
            const uint32_t col = rw_candidate(a, i0, k, cell, ai, j);
            act = !a.symmetric || col <= (uint32_t)r;
            if (act) kk = rw_find(S, nl, col);
        }
        const int at = rw_place(run, act, kk, lane);
        if (act) pos[(size_t)rw_slot(a, ai, j) * a.n_cells + cell] = c0 + at;
    }
}

// mirror position of every stored entry (r, c), c < r: the place of column r in row c of the full pattern
__global__ void k_mirror_pos(int64_t nu, int shift, const uint64_t* __restrict__ ukeys, const int32_t* __restrict__ rowptr,
                             const int32_t* __restrict__ colidx, int32_t* __restrict__ dst_b) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nu) return;
    const uint64_t key = ukeys[t];
    const int r = (int)(key >> shift), c = (int)(key & ((uint64_t(1) << shift) - 1));
    int f = -1;
    if (r != c) {
        int lo = rowptr[c], hi = rowptr[c + 1];
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (colidx[mid] < r) lo = mid + 1;
            else hi = mid;
        }
        f = lo;
    }
    dst_b[t] = f;
}

static int exclusive_scan_inplace(int32_t* p, int count, cudaStream_t st) {
    size_t tb = 0;
    FDB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, p, p, count, st));
    DevBuf<char> tmp;
    FDB_TRY(tmp.alloc(tb));
    FDB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, p, p, count, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    return FDB_OK;
}

// forward declarations of the incidence kernels (defined with the load-vector map below)
__global__ void k_emit_dof_keys(int64_t total, const int32_t* __restrict__ dofs_soa, int n_cells, int nb,
                                uint32_t* __restrict__ keys, uint32_t* __restrict__ ids);
__global__ void k_dof_seg(int n_dofs, int64_t total, const uint32_t* __restrict__ keys, int32_t* __restrict__ seg);

static int build_pattern_rows(fdb_space* s, int symmetric, bool* done) {
    *done = false;
    Pattern& P = s->pat[symmetric ? 1 : 0];
    cudaStream_t st = s->stream;
    const int nb = s->nb, n_cells = s->n_cells, n = s->n_dofs, B = 256;
    const int64_t total = (int64_t)n_cells * nb, nc = P.n_contrib;
    const int shift = bits_for(n);
    DevBuf<uint32_t> inc;
    DevBuf<int32_t> inc_ptr;
    FDB_TRY(inc_ptr.alloc((size_t)n + 1));
    {
        DevBuf<uint32_t> k0, k1, v0;
        FDB_TRY(k0.alloc(total)); FDB_TRY(k1.alloc(total)); FDB_TRY(v0.alloc(total)); FDB_TRY(inc.alloc(total));
        k_emit_dof_keys<<<grid_for(total, B), B, 0, st>>>(total, s->dofs.p, n_cells, nb, k0.p, v0.p);
        FDB_CUDA(cudaGetLastError());
        FDB_TRY(radix_sort_pairs(k0, k1, v0, inc, total, shift, st));   // stable: ascending cells inside a dof
        k_dof_seg<<<grid_for(n + 1, B), B, 0, st>>>(n, total, k1.p, inc_ptr.p);
        FDB_CUDA(cudaGetLastError());
        FDB_CUDA(cudaStreamSynchronize(st));
    }
    RowArgs a{nb, n_cells, symmetric ? 1 : 0, s->dofs.p, inc.p, inc_ptr.p};
    DevBuf<int32_t> lptr, cptr;
    DevBuf<int> flag;
    FDB_TRY(P.rowptr.alloc((size_t)n + 1)); FDB_TRY(lptr.alloc((size_t)n + 1)); FDB_TRY(cptr.alloc((size_t)n + 1));
    FDB_TRY(flag.alloc(1));
    FDB_CUDA(cudaMemsetAsync(P.rowptr.p + n, 0, sizeof(int32_t), st));
    FDB_CUDA(cudaMemsetAsync(lptr.p + n, 0, sizeof(int32_t), st));
    FDB_CUDA(cudaMemsetAsync(cptr.p + n, 0, sizeof(int32_t), st));
    FDB_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), st));
    const unsigned grid = grid_for(n, RW_WARPS);
    DevBuf<uint32_t> scratch;   // sorted distinct columns of every row, at the offset of the row's first emitted candidate
    FDB_TRY(scratch.alloc((size_t)total * nb));
    k_row_counts<<<grid, 32 * RW_WARPS, 0, st>>>(n, a, P.rowptr.p, lptr.p, cptr.p, scratch.p, flag.p);
    FDB_CUDA(cudaGetLastError());
    int overflow = 0;
    FDB_CUDA(cudaMemcpyAsync(&overflow, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    if (overflow) return FDB_OK;   // a row with too many distinct columns: the sort path handles any mesh
    FDB_TRY(exclusive_scan_inplace(P.rowptr.p, n + 1, st));
    FDB_TRY(exclusive_scan_inplace(lptr.p, n + 1, st));
    FDB_TRY(exclusive_scan_inplace(cptr.p, n + 1, st));
    int32_t tot[3] = {0, 0, 0};
    FDB_CUDA(cudaMemcpyAsync(&tot[0], P.rowptr.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(&tot[1], lptr.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(&tot[2], cptr.p + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    FDB_CHECK((int64_t)tot[2] == nc, FDB_ERR_STATE, "row-wise pattern build: the rows do not account for every emitted triplet");
    P.nnz = tot[0];
    P.n_unique = tot[1];
    FDB_TRY(P.colidx.alloc(P.nnz)); FDB_TRY(P.seg.alloc(P.n_unique + 1)); FDB_TRY(P.ukeys.alloc(P.n_unique));
    FDB_TRY(P.pos.alloc(nc)); FDB_TRY(P.dst_a.alloc(P.n_unique));
    k_row_fill<<<grid, 32 * RW_WARPS, 0, st>>>(n, a, shift, P.rowptr.p, lptr.p, cptr.p, scratch.p, P.colidx.p, P.ukeys.p,
                                              P.dst_a.p, P.seg.p, P.pos.p);
    FDB_CUDA(cudaGetLastError());
    if (symmetric) {
        FDB_TRY(P.dst_b.alloc(P.n_unique));
        k_mirror_pos<<<grid_for(P.n_unique, B), B, 0, st>>>(P.n_unique, shift, P.ukeys.p, P.rowptr.p, P.colidx.p, P.dst_b.p);
        FDB_CUDA(cudaGetLastError());
    }
    FDB_TRY(P.diag.alloc(n));
    k_diag<<<grid_for(n, B), B, 0, st>>>(n, P.rowptr.p, P.colidx.p, P.diag.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(st));
    P.shift = shift;
    *done = true;
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
    if (!getenv("FDB_PATTERN_SORT")) {   // row-wise build (no sort of the emitted triplets); FDB_PATTERN_SORT=1: sort path
        bool done = false;
        FDB_TRY(build_pattern_rows(s, symmetric, &done));
        if (done) {
            P.built = true;
            if (getenv("FDB_FUSED_EAGER")) FDB_TRY(ensure_fused_plan(s, &P));
            return FDB_OK;
        }
    }

    DevBuf<uint64_t> k0, k1;
    DevBuf<uint64_t>& ukeys = P.ukeys;  // kept: the fused plan is built lazily from (ukeys, seg, pos)
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
    P.shift = shift;

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
    if (getenv("FDB_FUSED_EAGER")) FDB_TRY(ensure_fused_plan(s, &P));
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
