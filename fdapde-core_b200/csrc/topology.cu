// K1: P2 degree-of-freedom enumeration on the device (SURVEY 8a rows A2 + A3).
//
// Replaces the reference's hash-map scans
//   Triangulation<2,N> ctor              geometry/triangulation.h:143-196  (edge id = first occurrence scanning cells
//                                        ascending x local pairs (0,1),(0,2),(1,2); boundary edge <=> one cell)
//   Triangulation<3,3> ctor              geometry/triangulation.h:319-399  (faces (0,1,2),(0,1,3),(0,2,3),(1,2,3); edges
//                                        numbered inside each NEW face from its sorted node triple; boundary edge <=>
//                                        both end nodes on the boundary, :376)
//   LagrangianBasis::enumerate_dofs      basis/lagrangian_basis.h:94-136   (dof = n_nodes + edge id at the local slot of
//                                        the edge's reference midpoint)
// by a stable radix sort: key = sorted node pair, value = scan position.  The first element of every key segment
// is the first occurrence; ranking the unique keys by that position reproduces the reference's edge ids exactly.
// P2 on tetrahedra does not exist in the reference; the slot convention is the extension A10.
#include <cub/cub.cuh>

#include "common.cuh"

namespace fdb {

static inline int bits_for(int64_t n) {
    int b = 1;
    while ((int64_t(1) << b) < n) ++b;
    return b;
}
static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// local vertex pair -> local dof slot.  2D: pairs (0,1),(0,2),(1,2) -> 3,4,5 (reference_element.h:60-62).
// 3D: midpoints of ReferenceElement<3,2>::nodes (reference_element.h:93-96).
__device__ __forceinline__ int slot_of_pair(int M, int a, int b) {  // a < b
    if (M == 2) return 3 + (a == 0 ? b - 1 : 2);
    const int code = a * 4 + b;
    switch (code) {
    case 1: return 6;   // (0,1)
    case 2: return 5;   // (0,2)
    case 3: return 9;   // (0,3)
    case 6: return 4;   // (1,2)
    case 7: return 7;   // (1,3)
    default: return 8;  // (2,3)
    }
}

// scan slot t -> (cell, local vertex pair), in the reference's scan order
__device__ __forceinline__ void slot_pair(int M, const int32_t* __restrict__ c, int j, int& la, int& lb) {
    if (M == 2) {
        la = (j == 2) ? 1 : 0;
        lb = (j == 0) ? 1 : 2;
        return;
    }
    // 3D: j = f*3 + k; face f = local vertices with index != 3-f..., sorted by global node id
    const int f = j / 3, k = j % 3;
    int lv[3];
    // faces (0,1,2),(0,1,3),(0,2,3),(1,2,3)
    lv[0] = (f == 3) ? 1 : 0;
    lv[1] = (f <= 1) ? 1 : 2;
    lv[2] = (f == 0) ? 2 : 3;
    // sort the three local vertices by node id
    if (c[lv[0]] > c[lv[1]]) { int t = lv[0]; lv[0] = lv[1]; lv[1] = t; }
    if (c[lv[1]] > c[lv[2]]) { int t = lv[1]; lv[1] = lv[2]; lv[2] = t; }
    if (c[lv[0]] > c[lv[1]]) { int t = lv[0]; lv[0] = lv[1]; lv[1] = t; }
    la = (k == 2) ? lv[1] : lv[0];
    lb = (k == 0) ? lv[1] : lv[2];
}

__global__ void k_edge_keys(int M, int n_cells, int per_cell, int shift, const int32_t* __restrict__ cells,
                            uint64_t* __restrict__ keys, uint32_t* __restrict__ ids) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_cells * per_cell) return;
    int e = (int)(t / per_cell), j = (int)(t % per_cell);
    const int32_t* c = cells + (size_t)e * (M + 1);
    int la, lb;
    slot_pair(M, c, j, la, lb);
    uint32_t a = (uint32_t)c[la], b = (uint32_t)c[lb];
    if (a > b) { uint32_t tmp = a; a = b; b = tmp; }
    keys[t] = ((uint64_t)a << shift) | b;
    ids[t] = (uint32_t)t;
}

__global__ void k_edge_heads(int64_t n, const uint64_t* __restrict__ keys, int32_t* __restrict__ flags) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    flags[t] = (t == 0 || keys[t] != keys[t - 1]) ? 1 : 0;
}

// per unique edge: first scan position (sort key for the ranking) and its own index
__global__ void k_edge_first(int64_t n, const uint32_t* __restrict__ ids, const int32_t* __restrict__ scan,
                             uint32_t* __restrict__ first_pos, uint32_t* __restrict__ uidx, int32_t* __restrict__ seg) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    bool head = (t == 0) || (scan[t] != scan[t - 1]);
    int32_t u = scan[t] - 1;
    if (head) {
        first_pos[u] = ids[t];
        uidx[u] = (uint32_t)u;
        seg[u] = (int32_t)t;
    }
    if (t == n - 1) seg[u + 1] = (int32_t)n;
}

__global__ void k_edge_rank(int n_edges, const uint32_t* __restrict__ uidx_sorted, int32_t* __restrict__ edge_id) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_edges) edge_id[uidx_sorted[r]] = r;
}

__global__ void k_edge_dofs(int M, int64_t n, int n_cells, int n_nodes, int per_cell, int shift,
                            const int32_t* __restrict__ cells, const uint64_t* __restrict__ keys,
                            const uint32_t* __restrict__ ids, const int32_t* __restrict__ scan,
                            const int32_t* __restrict__ seg, const int32_t* __restrict__ edge_id,
                            const uint8_t* __restrict__ bnodes, int32_t* __restrict__ dofs, uint8_t* __restrict__ bdofs) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    int32_t u = scan[t] - 1;
    int id = edge_id[u];
    uint32_t pos = ids[t];
    int e = (int)(pos / per_cell), j = (int)(pos % per_cell);
    const int32_t* c = cells + (size_t)e * (M + 1);
    int la, lb;
    slot_pair(M, c, j, la, lb);
    if (la > lb) { int tmp = la; la = lb; lb = tmp; }
    dofs[(size_t)slot_of_pair(M, la, lb) * n_cells + e] = n_nodes + id;
    if (seg[u] == (int32_t)t) {  // once per edge
        uint8_t flag;
        if (M == 2) flag = (seg[u + 1] - seg[u] == 1);
        else {
            uint64_t mask = (uint64_t(1) << shift) - 1;
            uint32_t a = (uint32_t)(keys[t] >> shift), b = (uint32_t)(keys[t] & mask);
            flag = bnodes ? (bnodes[a] && bnodes[b]) : 0;
        }
        bdofs[n_nodes + id] = flag;
    }
}

__global__ void k_vertex_dofs(int M, int n_cells, const int32_t* __restrict__ cells, int32_t* __restrict__ dofs) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_cells * (M + 1)) return;
    int e = (int)(t / (M + 1)), k = (int)(t % (M + 1));
    dofs[(size_t)k * n_cells + e] = cells[t];
}

template <typename K, typename V>
static int sort_pairs(DevBuf<K>& k_in, DevBuf<K>& k_out, DevBuf<V>& v_in, DevBuf<V>& v_out, int64_t n, int end_bit) {
    size_t tmp_bytes = 0;
    FDB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, (int)n, 0, end_bit));
    DevBuf<char> tmp;
    FDB_TRY(tmp.alloc(tmp_bytes));
    FDB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, (int)n, 0, end_bit));
    FDB_CUDA(cudaDeviceSynchronize());
    return FDB_OK;
}

int enumerate_dofs(int M, int R, int n_nodes, int n_cells, const int32_t* cells_h, const uint8_t* bnodes_h,
                   int32_t* dofs_h, uint8_t* bdofs_h, int* n_dofs_out) {
    FDB_CHECK((M == 2 || M == 3) && (R == 1 || R == 2), FDB_ERR_UNSUPPORTED, "only M in {2,3}, R in {1,2}");
    FDB_CHECK(cells_h && dofs_h && bdofs_h && n_dofs_out && n_nodes > 0 && n_cells > 0, FDB_ERR_ARG, "bad argument");
    const int nv = M + 1, nb = (R == 1) ? nv : nv * (nv + 1) / 2;
    const int B = 256;
    DevBuf<int32_t> cells, dofs;
    DevBuf<uint8_t> bnodes;
    FDB_TRY(cells.alloc((size_t)n_cells * nv));
    FDB_TRY(dofs.alloc((size_t)n_cells * nb));
    FDB_CUDA(cudaMemcpy(cells.p, cells_h, sizeof(int32_t) * (size_t)n_cells * nv, cudaMemcpyHostToDevice));
    if (bnodes_h) {
        FDB_TRY(bnodes.alloc(n_nodes));
        FDB_CUDA(cudaMemcpy(bnodes.p, bnodes_h, n_nodes, cudaMemcpyHostToDevice));
    }
    k_vertex_dofs<<<grid_for((int64_t)n_cells * nv, B), B>>>(M, n_cells, cells.p, dofs.p);
    FDB_CUDA(cudaGetLastError());
    if (bnodes_h) memcpy(bdofs_h, bnodes_h, n_nodes);
    else memset(bdofs_h, 0, n_nodes);
    if (R == 1) {
        FDB_CUDA(cudaMemcpy(dofs_h, dofs.p, sizeof(int32_t) * (size_t)n_cells * nb, cudaMemcpyDeviceToHost));
        *n_dofs_out = n_nodes;
        return FDB_OK;
    }
    const int per_cell = (M == 2) ? 3 : 12;
    const int64_t ns = (int64_t)n_cells * per_cell;
    FDB_CHECK(ns < (int64_t(1) << 31), FDB_ERR_UNSUPPORTED, "too many cells for one device pass");
    const int shift = bits_for(n_nodes);
    DevBuf<uint64_t> k0, k1;
    DevBuf<uint32_t> v0, v1;
    DevBuf<int32_t> scan;
    FDB_TRY(k0.alloc(ns)); FDB_TRY(k1.alloc(ns)); FDB_TRY(v0.alloc(ns)); FDB_TRY(v1.alloc(ns)); FDB_TRY(scan.alloc(ns));
    k_edge_keys<<<grid_for(ns, B), B>>>(M, n_cells, per_cell, shift, cells.p, k0.p, v0.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(sort_pairs(k0, k1, v0, v1, ns, 2 * shift));
    k_edge_heads<<<grid_for(ns, B), B>>>(ns, k1.p, scan.p);
    FDB_CUDA(cudaGetLastError());
    {
        size_t tmp_bytes = 0;
        FDB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, scan.p, scan.p, (int)ns));
        DevBuf<char> tmp;
        FDB_TRY(tmp.alloc(tmp_bytes));
        FDB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tmp_bytes, scan.p, scan.p, (int)ns));
        FDB_CUDA(cudaDeviceSynchronize());
    }
    int32_t n_edges = 0;
    FDB_CUDA(cudaMemcpy(&n_edges, scan.p + (ns - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
    DevBuf<uint32_t> fp0, fp1, ui0, ui1;
    DevBuf<int32_t> seg, edge_id;
    FDB_TRY(fp0.alloc(n_edges)); FDB_TRY(fp1.alloc(n_edges)); FDB_TRY(ui0.alloc(n_edges)); FDB_TRY(ui1.alloc(n_edges));
    FDB_TRY(seg.alloc((size_t)n_edges + 1));
    FDB_TRY(edge_id.alloc(n_edges));
    k_edge_first<<<grid_for(ns, B), B>>>(ns, v1.p, scan.p, fp0.p, ui0.p, seg.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(sort_pairs(fp0, fp1, ui0, ui1, n_edges, bits_for(ns)));
    k_edge_rank<<<grid_for(n_edges, B), B>>>(n_edges, ui1.p, edge_id.p);
    FDB_CUDA(cudaGetLastError());
    DevBuf<uint8_t> bdofs;
    FDB_TRY(bdofs.alloc((size_t)n_nodes + n_edges));
    k_edge_dofs<<<grid_for(ns, B), B>>>(M, ns, n_cells, n_nodes, per_cell, shift, cells.p, k1.p, v1.p, scan.p, seg.p,
                                        edge_id.p, bnodes_h ? bnodes.p : nullptr, dofs.p, bdofs.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaDeviceSynchronize());
    FDB_CUDA(cudaMemcpy(dofs_h, dofs.p, sizeof(int32_t) * (size_t)n_cells * nb, cudaMemcpyDeviceToHost));
    FDB_CUDA(cudaMemcpy(bdofs_h + n_nodes, bdofs.p + n_nodes, n_edges, cudaMemcpyDeviceToHost));
    *n_dofs_out = n_nodes + n_edges;
    return FDB_OK;
}

}  // namespace fdb
