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

// =====================================================================================================================
// Full mesh topology on the device: what the reference's Triangulation constructors build with
// std::unordered_map scans (geometry/triangulation.h:143-196 for triangles, :319-399 for tetrahedra) --
//   facets (edges of triangles / faces of tetrahedra) numbered by first occurrence, cell -> facets, facet -> cells,
//   neighbours (column j = cell across the facet opposite to local vertex j, -1 on the boundary), boundary facets,
//   and in 3D the edges numbered inside every new face, face -> edges, boundary edges, edge -> cells --
// here by stable radix sorts: key = sorted node tuple, value = scan slot.  The first element of a key segment is the
// first occurrence (stable sort of ascending slots), ranking the unique keys by it reproduces the reference's ids, and
// the two elements of a segment are the two cells that share the facet.
namespace fdb {

struct Topology {
    int M = 0, n_nodes = 0, n_cells = 0, n_facets = 0, n_edges = 0;
    int64_t n_edge_cells = 0;
    DevBuf<int32_t> neighbors, facets, cell_to_facets, facet_to_cells, edges, face_to_edges, edge_cell_ptr, edge_cells;
    DevBuf<uint8_t> facet_boundary, edge_boundary;
};

// local vertices of facet j of a cell, combinations<M, M+1>() order; the vertex opposite to facet j is M - j
__device__ __forceinline__ void facet_nodes(int M, const int32_t* __restrict__ c, int j, int32_t* f) {
    if (M == 2) {
        f[0] = c[(j == 2) ? 1 : 0];
        f[1] = c[(j == 0) ? 1 : 2];
        if (f[0] > f[1]) { int32_t t = f[0]; f[0] = f[1]; f[1] = t; }
    } else {
        f[0] = c[(j == 3) ? 1 : 0];
        f[1] = c[(j <= 1) ? 1 : 2];
        f[2] = c[(j == 0) ? 2 : 3];
        if (f[0] > f[1]) { int32_t t = f[0]; f[0] = f[1]; f[1] = t; }
        if (f[1] > f[2]) { int32_t t = f[1]; f[1] = f[2]; f[2] = t; }
        if (f[0] > f[1]) { int32_t t = f[0]; f[0] = f[1]; f[1] = t; }
    }
}

// pass keys of the facet slots: low = last node (3D) / nothing (2D), high = (first node << 32) | second node
__global__ void k_facet_keys(int M, int64_t ns, const int32_t* __restrict__ cells, uint32_t* __restrict__ klow,
                             uint64_t* __restrict__ khigh, uint32_t* __restrict__ ids) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ns) return;
    const int nv = M + 1;
    int32_t f[3];
    facet_nodes(M, cells + (t / nv) * nv, (int)(t % nv), f);
    khigh[t] = ((uint64_t)(uint32_t)f[0] << 32) | (uint32_t)f[1];
    if (M == 3) klow[t] = (uint32_t)f[2];
    ids[t] = (uint32_t)t;
}
__global__ void k_gather_u64(int64_t n, const uint32_t* __restrict__ idx, const uint64_t* __restrict__ in,
                             uint64_t* __restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = in[idx[t]];
}
// head flags of the sorted facet slots (all nodes of the tuple compared)
__global__ void k_facet_heads(int M, int64_t ns, const int32_t* __restrict__ cells, const uint32_t* __restrict__ ids,
                              int32_t* __restrict__ flags) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ns) return;
    const int nv = M + 1;
    int head = 1;
    if (t > 0) {
        int32_t a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
        const uint32_t p = ids[t], q = ids[t - 1];
        facet_nodes(M, cells + (size_t)(p / nv) * nv, (int)(p % nv), a);
        facet_nodes(M, cells + (size_t)(q / nv) * nv, (int)(q % nv), b);
        head = (a[0] != b[0]) || (a[1] != b[1]) || (M == 3 && a[2] != b[2]);
    }
    flags[t] = head;
}
__global__ void k_facet_first(int64_t ns, const uint32_t* __restrict__ ids, const int32_t* __restrict__ scan,
                              uint32_t* __restrict__ first_pos, uint32_t* __restrict__ uidx) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ns) return;
    if (t == 0 || scan[t] != scan[t - 1]) {
        const int32_t u = scan[t] - 1;
        first_pos[u] = ids[t];
        uidx[u] = (uint32_t)u;
    }
}
__global__ void k_fill_i32(int64_t n, int32_t v, int32_t* __restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = v;
}
// one thread per sorted slot: facet id, cell -> facet, facet -> cells, neighbours, boundary marker
__global__ void k_facet_link(int M, int64_t ns, const int32_t* __restrict__ cells, const uint32_t* __restrict__ ids,
                             const int32_t* __restrict__ scan, const int32_t* __restrict__ facet_id,
                             int32_t* __restrict__ facets, int32_t* __restrict__ cell_to_facets,
                             int32_t* __restrict__ facet_to_cells, uint8_t* __restrict__ facet_boundary,
                             int32_t* __restrict__ neighbors) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ns) return;
    const int nv = M + 1;
    const int32_t u = scan[t] - 1;
    const int fid = facet_id[u];
    const uint32_t p = ids[t];
    const int cell = (int)(p / nv), j = (int)(p % nv);
    cell_to_facets[p] = fid;
    const bool head = (t == 0) || (scan[t - 1] != scan[t]);
    const bool has_next = (t + 1 < ns) && (scan[t + 1] == scan[t]);
    if (head) {
        int32_t f[3];
        facet_nodes(M, cells + (size_t)cell * nv, j, f);
        for (int k = 0; k < M; ++k) facets[(size_t)fid * M + k] = f[k];
        facet_to_cells[2 * (size_t)fid] = cell;
        facet_boundary[fid] = has_next ? 0 : 1;
        if (has_next) {   // the second cell of the facet (a facet of a manifold mesh has at most two)
            const uint32_t q = ids[t + 1];
            const int other = (int)(q / nv), jo = (int)(q % nv);
            facet_to_cells[2 * (size_t)fid + 1] = other;
            neighbors[(size_t)cell * nv + (M - j)] = other;     // vertex opposite to facet j is M - j
            neighbors[(size_t)other * nv + (M - jo)] = cell;
        }
    }
}
// 3D: edge slots of the faces in face-id order, pairs (0,1),(0,2),(1,2) of the sorted face triple
__global__ void k_face_edge_keys(int64_t ns, const int32_t* __restrict__ faces, uint64_t* __restrict__ keys,
                                 uint32_t* __restrict__ ids) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ns) return;
    const int32_t* f = faces + (t / 3) * 3;
    const int k = (int)(t % 3);
    const uint32_t a = (uint32_t)f[(k == 2) ? 1 : 0], b = (uint32_t)f[(k == 0) ? 1 : 2];
    keys[t] = ((uint64_t)a << 32) | b;   // the triple is sorted: a < b
    ids[t] = (uint32_t)t;
}
__global__ void k_key_heads(int64_t n, const uint64_t* __restrict__ keys, int32_t* __restrict__ flags) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) flags[t] = (t == 0 || keys[t] != keys[t - 1]) ? 1 : 0;
}
__global__ void k_edge_link(int64_t ns, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ ids,
                            const int32_t* __restrict__ scan, const int32_t* __restrict__ edge_id,
                            const uint8_t* __restrict__ bnodes, int32_t* __restrict__ edges,
                            int32_t* __restrict__ face_to_edges, uint8_t* __restrict__ edge_boundary,
                            uint64_t* __restrict__ ukeys, int32_t* __restrict__ ukey_id) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= ns) return;
    const int32_t u = scan[t] - 1;
    const int eid = edge_id[u];
    face_to_edges[ids[t]] = eid;
    if (t == 0 || scan[t - 1] != scan[t]) {
        const uint32_t a = (uint32_t)(keys[t] >> 32), b = (uint32_t)(keys[t] & 0xffffffffu);
        edges[2 * (size_t)eid] = (int32_t)a;
        edges[2 * (size_t)eid + 1] = (int32_t)b;
        edge_boundary[eid] = bnodes ? (bnodes[a] && bnodes[b]) : 0;
        ukeys[u] = keys[t];     // unique keys ascending: the lookup table of the (edge, cell) pass
        ukey_id[u] = eid;
    }
}
// 3D: (edge id << 32 | cell) for the six edges of every cell
__global__ void k_cell_edge_pairs(int n_cells, int n_ukeys, const int32_t* __restrict__ cells,
                                  const uint64_t* __restrict__ ukeys, const int32_t* __restrict__ ukey_id,
                                  uint64_t* __restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_cells * 6) return;
    const int cell = (int)(t / 6), k = (int)(t % 6);
    const int la = (k < 3) ? 0 : (k < 5 ? 1 : 2), lb = (k < 3) ? k + 1 : (k < 5 ? k - 1 : 3);
    uint32_t a = (uint32_t)cells[(size_t)cell * 4 + la], b = (uint32_t)cells[(size_t)cell * 4 + lb];
    if (a > b) { uint32_t tmp = a; a = b; b = tmp; }
    const uint64_t key = ((uint64_t)a << 32) | b;
    int lo = 0, hi = n_ukeys;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ukeys[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    out[t] = ((uint64_t)(uint32_t)ukey_id[lo] << 32) | (uint32_t)cell;
}
__global__ void k_edge_cell_lists(int64_t n, int n_edges, const uint64_t* __restrict__ sorted, int32_t* __restrict__ ptr,
                                  int32_t* __restrict__ list) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int e = (int)(sorted[t] >> 32);
    list[t] = (int32_t)(sorted[t] & 0xffffffffu);
    const int prev = (t == 0) ? -1 : (int)(sorted[t - 1] >> 32);
    for (int k = prev + 1; k <= e; ++k) ptr[k] = (int32_t)t;   // every edge has at least one cell, so this is one store
    if (t == n - 1) ptr[n_edges] = (int32_t)n;
}

static int inclusive_scan(DevBuf<int32_t>& v, int64_t n) {
    size_t tb = 0;
    FDB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, v.p, v.p, (int)n));
    DevBuf<char> tmp;
    FDB_TRY(tmp.alloc(tb));
    FDB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, v.p, v.p, (int)n));
    FDB_CUDA(cudaDeviceSynchronize());
    return FDB_OK;
}
static int sort_keys64(DevBuf<uint64_t>& k_in, DevBuf<uint64_t>& k_out, int64_t n, int end_bit) {
    size_t tb = 0;
    FDB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, k_in.p, k_out.p, (int)n, 0, end_bit));
    DevBuf<char> tmp;
    FDB_TRY(tmp.alloc(tb));
    FDB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, k_in.p, k_out.p, (int)n, 0, end_bit));
    FDB_CUDA(cudaDeviceSynchronize());
    return FDB_OK;
}

// ranks the unique keys by the scan slot of their first occurrence: rank_of[u] = id the reference gives
static int rank_by_first(DevBuf<uint32_t>& first_pos, DevBuf<uint32_t>& uidx, int n_unique, int64_t ns, DevBuf<int32_t>& rank_of) {
    DevBuf<uint32_t> fp1, ui1;
    FDB_TRY(fp1.alloc(n_unique)); FDB_TRY(ui1.alloc(n_unique));
    FDB_TRY(sort_pairs(first_pos, fp1, uidx, ui1, n_unique, bits_for(ns)));
    FDB_TRY(rank_of.alloc(n_unique));
    k_edge_rank<<<grid_for(n_unique, 256), 256>>>(n_unique, ui1.p, rank_of.p);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

int build_topology(int M, int n_nodes, int n_cells, const int32_t* cells_h, const uint8_t* bnodes_h, Topology* T) {
    FDB_CHECK(M == 2 || M == 3, FDB_ERR_UNSUPPORTED, "mesh topology: triangles (M = 2) or tetrahedra (M = 3)");
    FDB_CHECK(cells_h && n_nodes > 0 && n_cells > 0 && T, FDB_ERR_ARG, "bad argument");
    const int nv = M + 1, B = 256;
    const int64_t ns = (int64_t)n_cells * nv;
    FDB_CHECK(ns * 3 < (int64_t(1) << 31), FDB_ERR_UNSUPPORTED, "too many cells for one device pass");
    T->M = M; T->n_nodes = n_nodes; T->n_cells = n_cells;
    DevBuf<int32_t> cells;
    DevBuf<uint8_t> bnodes;
    FDB_TRY(cells.alloc(ns));
    FDB_CUDA(cudaMemcpy(cells.p, cells_h, sizeof(int32_t) * ns, cudaMemcpyHostToDevice));
    if (bnodes_h) {
        FDB_TRY(bnodes.alloc(n_nodes));
        FDB_CUDA(cudaMemcpy(bnodes.p, bnodes_h, n_nodes, cudaMemcpyHostToDevice));
    }
    const int nbits = bits_for(n_nodes);
    // ---- facets: LSD over the node tuple -- (3D: last node first), then (first node, second node) ---------------------------
    DevBuf<uint32_t> kl0, kl1, id0, id1;
    DevBuf<uint64_t> kh0, kh1;
    FDB_TRY(kh0.alloc(ns)); FDB_TRY(kh1.alloc(ns)); FDB_TRY(id0.alloc(ns)); FDB_TRY(id1.alloc(ns));
    if (M == 3) { FDB_TRY(kl0.alloc(ns)); FDB_TRY(kl1.alloc(ns)); }
    k_facet_keys<<<grid_for(ns, B), B>>>(M, ns, cells.p, kl0.p, kh0.p, id0.p);
    FDB_CUDA(cudaGetLastError());
    if (M == 3) {
        FDB_TRY(sort_pairs(kl0, kl1, id0, id1, ns, nbits));                 // by the last node
        k_gather_u64<<<grid_for(ns, B), B>>>(ns, id1.p, kh0.p, kh1.p);      // high keys in that order
        FDB_CUDA(cudaGetLastError());
        FDB_TRY(sort_pairs(kh1, kh0, id1, id0, ns, 32 + nbits));            // stable: by (first, second) node
        std::swap(id0.p, id1.p);                                            // sorted slots -> id1
        std::swap(id0.n, id1.n);
    } else {
        FDB_TRY(sort_pairs(kh0, kh1, id0, id1, ns, 32 + nbits));
    }
    DevBuf<int32_t> scan;
    FDB_TRY(scan.alloc(ns));
    k_facet_heads<<<grid_for(ns, B), B>>>(M, ns, cells.p, id1.p, scan.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(inclusive_scan(scan, ns));
    int32_t nf = 0;
    FDB_CUDA(cudaMemcpy(&nf, scan.p + (ns - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
    T->n_facets = nf;
    DevBuf<uint32_t> fp, ui;
    DevBuf<int32_t> facet_id;
    FDB_TRY(fp.alloc(nf)); FDB_TRY(ui.alloc(nf));
    k_facet_first<<<grid_for(ns, B), B>>>(ns, id1.p, scan.p, fp.p, ui.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(rank_by_first(fp, ui, nf, ns, facet_id));
    FDB_TRY(T->neighbors.alloc(ns)); FDB_TRY(T->cell_to_facets.alloc(ns));
    FDB_TRY(T->facets.alloc((size_t)nf * M)); FDB_TRY(T->facet_to_cells.alloc((size_t)nf * 2));
    FDB_TRY(T->facet_boundary.alloc(nf));
    k_fill_i32<<<grid_for(ns, B), B>>>(ns, -1, T->neighbors.p);
    k_fill_i32<<<grid_for((int64_t)nf * 2, B), B>>>((int64_t)nf * 2, -1, T->facet_to_cells.p);
    k_facet_link<<<grid_for(ns, B), B>>>(M, ns, cells.p, id1.p, scan.p, facet_id.p, T->facets.p, T->cell_to_facets.p,
                                         T->facet_to_cells.p, T->facet_boundary.p, T->neighbors.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaDeviceSynchronize());
    if (M == 2) { T->n_edges = nf; return FDB_OK; }
    // ---- 3D: edges numbered inside the faces (face-id order), face -> edges, boundary edges, edge -> cells ---------------
    const int64_t nes = (int64_t)nf * 3;
    DevBuf<uint64_t> ek0, ek1;
    DevBuf<uint32_t> ei0, ei1;
    FDB_TRY(ek0.alloc(nes)); FDB_TRY(ek1.alloc(nes)); FDB_TRY(ei0.alloc(nes)); FDB_TRY(ei1.alloc(nes));
    k_face_edge_keys<<<grid_for(nes, B), B>>>(nes, T->facets.p, ek0.p, ei0.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(sort_pairs(ek0, ek1, ei0, ei1, nes, 32 + nbits));
    DevBuf<int32_t> escan;
    FDB_TRY(escan.alloc(nes));
    k_key_heads<<<grid_for(nes, B), B>>>(nes, ek1.p, escan.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(inclusive_scan(escan, nes));
    int32_t ne = 0;
    FDB_CUDA(cudaMemcpy(&ne, escan.p + (nes - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
    T->n_edges = ne;
    DevBuf<uint32_t> efp, eui;
    DevBuf<int32_t> edge_id, ukey_id;
    DevBuf<uint64_t> ukeys;
    FDB_TRY(efp.alloc(ne)); FDB_TRY(eui.alloc(ne)); FDB_TRY(ukeys.alloc(ne)); FDB_TRY(ukey_id.alloc(ne));
    k_facet_first<<<grid_for(nes, B), B>>>(nes, ei1.p, escan.p, efp.p, eui.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(rank_by_first(efp, eui, ne, nes, edge_id));
    FDB_TRY(T->edges.alloc((size_t)ne * 2)); FDB_TRY(T->face_to_edges.alloc(nes)); FDB_TRY(T->edge_boundary.alloc(ne));
    k_edge_link<<<grid_for(nes, B), B>>>(nes, ek1.p, ei1.p, escan.p, edge_id.p, bnodes_h ? bnodes.p : nullptr, T->edges.p,
                                         T->face_to_edges.p, T->edge_boundary.p, ukeys.p, ukey_id.p);
    FDB_CUDA(cudaGetLastError());
    const int64_t np = (int64_t)n_cells * 6;
    DevBuf<uint64_t> pk0, pk1;
    FDB_TRY(pk0.alloc(np)); FDB_TRY(pk1.alloc(np));
    k_cell_edge_pairs<<<grid_for(np, B), B>>>(n_cells, ne, cells.p, ukeys.p, ukey_id.p, pk0.p);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(sort_keys64(pk0, pk1, np, 32 + bits_for(ne)));
    FDB_TRY(T->edge_cell_ptr.alloc((size_t)ne + 1)); FDB_TRY(T->edge_cells.alloc(np));
    k_edge_cell_lists<<<grid_for(np, B), B>>>(np, ne, pk1.p, T->edge_cell_ptr.p, T->edge_cells.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaDeviceSynchronize());
    T->n_edge_cells = np;
    return FDB_OK;
}

}  // namespace fdb

struct fdb_topology { fdb::Topology t; };

extern "C" {

int fdb_topology_create(fdb_topology** out, int M, int n_nodes, int n_cells, const int32_t* cells_rowmajor,
                        const uint8_t* boundary_nodes) {
    FDB_CHECK(out, FDB_ERR_ARG, "null output handle");
    *out = nullptr;
    fdb_topology* h = new fdb_topology();
    int rc = fdb::build_topology(M, n_nodes, n_cells, cells_rowmajor, boundary_nodes, &h->t);
    if (rc != FDB_OK) { delete h; return rc; }
    *out = h;
    return FDB_OK;
}
void fdb_topology_destroy(fdb_topology* h) { delete h; }

int fdb_topology_sizes(const fdb_topology* h, int* n_facets, int* n_edges, int64_t* n_edge_cells) {
    FDB_CHECK(h, FDB_ERR_ARG, "null topology");
    if (n_facets) *n_facets = h->t.n_facets;
    if (n_edges) *n_edges = h->t.n_edges;
    if (n_edge_cells) *n_edge_cells = h->t.n_edge_cells;
    return FDB_OK;
}

int fdb_topology_download(const fdb_topology* h, int32_t* neighbors, int32_t* facets, int32_t* cell_to_facets,
                          int32_t* facet_to_cells, uint8_t* facet_boundary, int32_t* edges, int32_t* face_to_edges,
                          uint8_t* edge_boundary, int32_t* edge_cell_ptr, int32_t* edge_cells) {
    FDB_CHECK(h, FDB_ERR_ARG, "null topology");
    const fdb::Topology& T = h->t;
    const size_t ns = (size_t)T.n_cells * (T.M + 1);
#define FDB_DL(dst, src, count, type) \
    if (dst && (count) > 0) FDB_CUDA(cudaMemcpy(dst, (src).p, sizeof(type) * (size_t)(count), cudaMemcpyDeviceToHost))
    FDB_DL(neighbors, T.neighbors, ns, int32_t);
    FDB_DL(facets, T.facets, (size_t)T.n_facets * T.M, int32_t);
    FDB_DL(cell_to_facets, T.cell_to_facets, ns, int32_t);
    FDB_DL(facet_to_cells, T.facet_to_cells, (size_t)T.n_facets * 2, int32_t);
    FDB_DL(facet_boundary, T.facet_boundary, T.n_facets, uint8_t);
    if (T.M == 3) {
        FDB_DL(edges, T.edges, (size_t)T.n_edges * 2, int32_t);
        FDB_DL(face_to_edges, T.face_to_edges, (size_t)T.n_facets * 3, int32_t);
        FDB_DL(edge_boundary, T.edge_boundary, T.n_edges, uint8_t);
        FDB_DL(edge_cell_ptr, T.edge_cell_ptr, (size_t)T.n_edges + 1, int32_t);
        FDB_DL(edge_cells, T.edge_cells, T.n_edge_cells, int32_t);
    }
#undef FDB_DL
    return FDB_OK;
}

}  // extern "C"
