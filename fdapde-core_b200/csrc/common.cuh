// Internal declarations shared by the translation units of libfdapde_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>
#include <vector>

#include "../../include/fdapde_b200.h"

namespace fdb {

void set_error(const std::string& msg);

#define FDB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            fdb::set_error(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                           std::to_string(__LINE__) + ")");                                         \
            return FDB_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define FDB_CHECK(cond, code, msg)         \
    do {                                   \
        if (!(cond)) {                     \
            fdb::set_error(msg);           \
            return code;                   \
        }                                  \
    } while (0)

#define FDB_TRY(expr)                \
    do {                             \
        int rc_ = (expr);            \
        if (rc_ != FDB_OK) return rc_; \
    } while (0)

// Device memory comes from a process-wide caching allocator (api.cu): released blocks are kept and reused, because
// cudaMalloc / cudaFree of the GB-sized temporaries of the pattern build cost far more than the kernels between them.
// Blocks are handed back to the driver by fdb_trim().  A block may be reused as soon as it is released, so a buffer
// must not be released while work that uses it is still queued on a stream other than the one that will reuse it
// (every build step synchronises its stream before its temporaries go out of scope).
void* pool_alloc(size_t bytes);
void pool_free(void* p);

// owning device buffer
template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) pool_free(p);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        p = static_cast<T*>(pool_alloc(count * sizeof(T)));
        if (!p) return FDB_ERR_CUDA;  // pool_alloc has set the error text
        n = count;
        return FDB_OK;
    }
};

struct LLWord { unsigned long long lo, hi; };  // 16 bytes per value: flag-in-data exchange word (solve_common.cuh)

constexpr int MAX_NB = 10;
constexpr int MAX_NQ = 6;
constexpr int MAX_D = 3;

// reference-element tables evaluated once on the host (A4/A5), staged through shared memory by the kernels
struct FeTables {
    int M, R, nb, nq;
    double w[MAX_NQ];                     // quadrature weights
    double qn[MAX_NQ * MAX_D];            // quadrature nodes (reference coordinates)
    double phi[MAX_NQ * MAX_NB];          // phi[q*nb + i]     = psi_i(p_q)
    double gref[MAX_NQ * MAX_NB * MAX_D]; // gref[(q*nb+i)*M+m] = d psi_i / d x_m (p_q)
    double refn[MAX_NB * MAX_D];          // reference nodes of the dofs
};
// polynomial form of the reference basis (point evaluation): psi_i(x) = sum_m coef[i*nb + m] * prod_d x_d^ex[m*M + d]
struct PolyTables {
    int M, R, nb, pad;
    int ex[MAX_NB * MAX_D];
    double coef[MAX_NB * MAX_NB];
};
int build_fe_tables(int M, int R, FeTables* t, PolyTables* poly = nullptr);

// sparsity pattern + scatter map of one symmetry class (K2)
struct Pattern {
    bool built = false;
    bool symmetric = false;
    int ne = 0;                 // emitted local entries per cell: nb(nb+1)/2 (symmetric) or nb*nb
    int64_t n_contrib = 0;      // n_cells * ne
    int64_t n_unique = 0;       // distinct (row,col) among the emitted triplets (lower triangle if symmetric)
    int64_t nnz = 0;            // stored entries of the full matrix
    DevBuf<int32_t> rowptr;     // n_dofs + 1   (== Eigen outer: the pattern is structurally symmetric)
    DevBuf<int32_t> colidx;     // nnz          (== Eigen inner)
    DevBuf<int32_t> pos;        // [ne][n_cells] slot of each local entry in the sorted contribution list
    DevBuf<int32_t> seg;        // n_unique + 1 segment offsets into the sorted contribution list
    DevBuf<int32_t> dst_a;      // n_unique     position of the entry in the full CSR arrays
    DevBuf<int32_t> dst_b;      // n_unique     position of its mirror (-1: diagonal / non-symmetric)
    DevBuf<int32_t> tperm;      // nnz          transpose permutation (CSC value k = CSR value tperm[k]); lazy
    DevBuf<int32_t> diag;       // n_dofs       position of the diagonal entry of each row (-1: none)
    DevBuf<int16_t> col16;      // nnz          colidx - row as 16 bits, when every offset fits (SpMV reads this instead)
    int col16_state = 0;        // 0 not tried, 1 usable, -1 does not fit
    // Sliced-ELL view for SpMV (slices of 32 rows, entries interleaved: step j of row 32 s + l at sell_ptr[s] + 32 j + l,
    // padded to the slice's longest row): one thread per row, every load of the value / column streams is a fully
    // coalesced line, and on banded matrices the x gathers of a step are (nearly) consecutive too.
    int sell_state = 0;         // 0 not tried, 1 built, -1 not worth it (padding too large)
    int64_t sell_slots = 0;
    DevBuf<int32_t> sell_ptr;   // n_slices + 1
    DevBuf<int32_t> sell_perm;  // slot -> position in the CSR arrays (-1: padding)
    DevBuf<int32_t> sell_col;   // slot -> column (padding: the row itself), or
    DevBuf<int16_t> sell_col16; // slot -> column - row when col16_state == 1

    // Fused plan (assembly without a materialised contribution list): rows are grouped into spatially compact
    // blocks (Morton order of a row's first incident cell); one CTA computes the local matrices of every cell
    // incident to its rows into shared memory and sums each stored entry of those rows from there, in the same
    // left-to-right emission order as the two-kernel path.
    DevBuf<uint64_t> ukeys;     // n_unique     (row << shift | col) of every stored entry (input of the fused plan)
    int shift = 0;
    int n_assemblies = 0;       // assemblies run on this pattern so far
    bool fused_tried = false;   // ensure_fused_plan has run
    bool fused = false;         // plan usable
    int f_rb = 0;               // rows per block
    int f_lcap = 0;             // shared-memory capacity in local entries (max over the blocks, padded)
    bool f_compact = false;     // compact records of the needed entries (P2) instead of slot-major whole local matrices (P1)
    int f_cells_cap = 0;        // slot-major layout: cells per slot row (max listed cells of a block, padded)
    int f_nblocks = 0;
    int f_threads = 256;        // threads per CTA of the fused kernel
    DevBuf<int32_t> f_rorder;   // n_dofs       rows in block order
    DevBuf<int32_t> f_urow;     // n_dofs + 1   row pointer into the unique-entry list
    DevBuf<int32_t> f_bcell_ptr;// nblocks + 1
    DevBuf<int32_t> f_bcells;   // cells of each block, in descending order of their slot masks
    DevBuf<unsigned long long> f_bmask;  // per listed cell: emission slots the block sums (bit s = slot s)
    DevBuf<uint16_t> f_bbase;   // per listed cell: start of its compact record in the block's shared-memory array
    DevBuf<uint16_t> f_lidx;    // n_contrib    shared-memory index (record start + rank of the slot) of every contribution,
                                //              block-major: block b owns [f_con_ptr[b], f_con_ptr[b+1])
    DevBuf<int32_t> f_bverts;   // vertex ids of the listed cells, (M+1) per cell, block-major
    DevBuf<int32_t> f_ent_ptr;  // nblocks + 1  stored entries before block b (block-major entry numbering)
    DevBuf<int32_t> f_con_ptr;  // nblocks + 1  contributions before block b
    DevBuf<int2> f_dst;         // n_unique     (position, mirror position or -1) of every entry, block-major (symmetric)
    DevBuf<int32_t> f_dst1;     // n_unique     position of every entry, block-major (non-symmetric patterns)
    bool f_dsm = false;         // destinations ride the bulk-copy prologue into shared memory (costs no resident CTA)
    DevBuf<uint16_t> f_segrel;  // n_unique + nblocks + 1: per block, entries + 1 segment offsets relative to the block
    DevBuf<int32_t> f_meta;     // per block: {first contribution, contributions, first entry, entries, first cell, cells, 0, 0}
    int f_max_ent = 0, f_max_con = 0;
    int f_max_cells = 0;        // most listed cells of a block
    // block-local node copies (P1): the distinct nodes of a block's cells are stored once per block, so their coordinates
    // ride the bulk-copy prologue into shared memory and phase 1 reads them there through 16-bit block-local indices
    bool f_nodes = false;
    int f_node_bytes = 0;       // shared memory for the coordinates of a block (max over the blocks, 16-byte granules)
    int f_node_z_off = 0;       // 3D: byte offset of the z array inside that region (x, y pairs come first)
    DevBuf<uint16_t> f_bvloc;   // 4 block-local node indices per listed cell (block-major)
    DevBuf<double> f_bcoords;   // (x, y) of every block's distinct nodes, block-major (16 bytes per node: with the z
                                // values in their own array the shared-memory copies use all 32 banks)
    DevBuf<double> f_bz;        // 3D: z of the same nodes
};

// dynamic shared memory of the fused kernel: local matrices + gather indices + segment offsets (+ destinations)
inline size_t fused_smem_bytes(const Pattern& P, bool dsm, int* con_cap_out = nullptr, int* ent_cap_out = nullptr) {
    const int con_cap = (P.f_max_con + 24) & ~7;   // room for the 16-byte alignment slack at both ends
    const int ent_cap = (P.f_max_ent + 24) & ~7;
    if (con_cap_out) *con_cap_out = con_cap;
    if (ent_cap_out) *ent_cap_out = ent_cap;
    const size_t dst_bytes = dsm ? ((size_t)(ent_cap + 8) * (P.symmetric ? 8 : 4) + 15) / 16 * 16 : 0;
    return sizeof(double) * (size_t)P.f_lcap + sizeof(uint16_t) * ((size_t)con_cap + ent_cap) + dst_bytes +
           ((P.f_nodes && !P.f_compact) ? (size_t)P.f_node_bytes : 0);   // P2: only the persistent kernel reads node copies
}
// byte offset of the block's node coordinates inside the dynamic shared memory (they come last)
inline size_t fused_coord_offset(const Pattern& P, bool dsm) {
    return fused_smem_bytes(P, dsm) - ((P.f_nodes && !P.f_compact) ? (size_t)P.f_node_bytes : 0);
}

// shared-memory layout of the persistent fused kernel (k_fused_persist): local matrices, then the single-buffered block
// lists (node indices of the cells, node coordinates, gather indices, segment offsets, destinations), 16-byte granules
struct PersistLayout { int off_ids, off_mask, off_base, off_coords, z_off, off_lidx, off_seg, off_dst, lcap_cells; };
inline size_t persist_layout(const Pattern& P, PersistLayout* L) {
    int con_cap, ent_cap;
    fused_smem_bytes(P, true, &con_cap, &ent_cap);
    auto up16 = [](size_t v) { return (v + 15) & ~size_t(15); };
    size_t off = sizeof(double) * (size_t)P.f_lcap;
    PersistLayout l;
    l.off_ids = (int)off;    off += up16((size_t)(P.f_max_cells + 4) * 8);
    l.off_mask = (int)off;   off += P.f_compact ? up16((size_t)(P.f_max_cells + 4) * 8) : 0;    // P2: needed-slot masks
    l.off_base = (int)off;   off += P.f_compact ? up16((size_t)(P.f_max_cells + 16) * 2) : 0;   //     record starts
    l.off_coords = (int)off; off += up16((size_t)P.f_node_bytes);
    l.z_off = P.f_node_z_off;
    l.off_lidx = (int)off;   off += up16(sizeof(uint16_t) * (size_t)con_cap);
    l.off_seg = (int)off;    off += up16(sizeof(uint16_t) * (size_t)ent_cap);
    l.off_dst = (int)off;    off += up16((size_t)(ent_cap + 8) * (P.symmetric ? 8 : 4));
    l.lcap_cells = P.f_cells_cap;
    if (L) *L = l;
    return off;
}
// two CTAs of the persistent kernel per SM (228 KB, 1 KB reserved per CTA)
inline bool persist_fits(const Pattern& P) { return P.f_nodes && 2 * (persist_layout(P, nullptr) + 1024) <= 228 * 1024; }

// per-dof gather lists for the load vector (K5)
struct ForcingMap {
    bool built = false;
    DevBuf<int32_t> pos;  // [nb][n_cells]
    DevBuf<int32_t> seg;  // n_dofs + 1
};

}  // namespace fdb

struct fdb_comm;
namespace fdb {
// row-block partition of one rank (multi-GPU solve): local dofs = [owned | halo grouped by neighbour rank]
// uniform-grid point locator (evaluate.cu): cells binned by bounding box
struct GridDesc {
    int g[3];                           // bins per axis (1 along unused axes)
    double lo[3], inv_h[3], eps[3];     // origin, 1 / bin size, inflation of the cell boxes
};
struct Locator {
    bool built = false;
    GridDesc grid;
    DevBuf<int32_t> bin_ptr, bin_cells; // cells of bin b: bin_cells[bin_ptr[b] .. bin_ptr[b+1])
};

struct Partition {
    fdb_comm* comm = nullptr;
    int n_owned = 0, n_send = 0, n_halo = 0;
    long long n_global = 0;                    // sum of n_owned over the ranks (rank-independent iteration budget)
    std::vector<int> nbr, send_off, recv_off;  // neighbour ranks; prefix offsets (size nbr + 1)
    DevBuf<int32_t> send_idx;                  // owned local indices to send, grouped by neighbour
    DevBuf<double> sendbuf;
    DevBuf<double> stage;                      // [0,16): per-rank sums, [16,32): all-reduced sums
    // peer-memory plan of the persistent solver (cudaIpc-mapped buffers of the other ranks), see comm.cu
    bool peer_ready = false;
    void* peer_buf = nullptr;                  // this rank's exported buffer: [halo LL words x2 | reduction LL words | error]
    std::vector<void*> peer_mapped;            // cudaIpcOpenMemHandle results (to close)
    void* peer_view = nullptr;                 // heap copy of the kernel's PeerView
    unsigned peer_tag = 0;                     // last exchange tag used (tags are never reused)
    ~Partition();
};
}  // namespace fdb

struct fdb_space {
    int M, N, R, nb, nq;
    int n_nodes, n_cells, n_dofs;
    fdb::FeTables tab_host;
    fdb::DevBuf<fdb::FeTables> tab;      // device copy
    fdb::DevBuf<double> tens;            // reference tensors of the constant-coefficient form: nb^2 rows of [T^mn | A^n | R | pad]
    fdb::PolyTables poly_host;
    fdb::DevBuf<fdb::PolyTables> poly;   // device copy
    fdb::Locator locator;                // built on the first point-location query
    fdb::DevBuf<double> coords;          // SoA [N][n_nodes]
    fdb::DevBuf<double> coords_pk;       // packed per node (3D: x y z pad, 2D: x y): one sector per gathered node
    int fused_threads = 0;               // > 0: override of the plan's threads per CTA (FDB_FUSED_THREADS)
    fdb::DevBuf<int32_t> verts;          // SoA [M+1][n_cells]  (aliases dofs when cells == NULL)
    const int32_t* verts_p = nullptr;
    fdb::DevBuf<int32_t> dofs;           // SoA [nb][n_cells]
    fdb::DevBuf<uint8_t> boundary;       // n_dofs
    bool has_boundary = false;
    bool dof0_rule = true;               // local dof 0 is always a Dirichlet dof (fem_solver_base.h:86)
    fdb::Pattern pat[2];                 // [0] general, [1] symmetric
    fdb::ForcingMap fmap;
    fdb::DevBuf<double> contrib;         // scratch: sorted contribution list (max over uses)
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool force_two_kernel = false;       // fdb_space_set_fused(s, 0): use the contribution-list path
    bool profile = false;                // per-kernel CUDA-event timing of the assembly (fdb_space_set_profiling)
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    bool ev_valid = false;
    int last_fused = -1;                 // path of the last assembly: 2 persistent fused kernel, 1 fused kernel, 0 contribution list + reduction
    bool last_persist = false;           // set by the persistent launch
    int last_launches = 0;               // kernels launched by the last assembly
    int device = 0;
    int sm_count = 148;
    int refs = 1;                        // owner + one per fdb_matrix
};

struct fdb_matrix {
    fdb_space* space = nullptr;
    const fdb::Pattern* pat = nullptr;  // set by the first assembly
    fdb::DevBuf<double> val;            // CSR(A) values
    bool assembled = false;
    uint64_t val_version = 0;           // bumped whenever the values change (assembly, Dirichlet rows, axpby)
    fdb::DevBuf<double> sell_val;       // values in the pattern's sliced-ELL order (refreshed lazily by SpMV)
    uint64_t sell_version = ~0ull;
    // solver workspace (lazily sized)
    fdb::DevBuf<double> work;
    fdb::DevBuf<double> partials;
    fdb::DevBuf<double> hist;
    fdb::Partition* part = nullptr;     // set by fdb_matrix_set_partition (owned)
    // persistent sliced-ELL solvers (solve_peer.cu)
    fdb::DevBuf<int32_t> slice_order;   // owned slices, interior first, halo-coupled last (partitioned matrices)
    int slice_order_n = -1;
    fdb::DevBuf<fdb::LLWord> local_red; // reduction lines of a single-GPU run
    unsigned local_tag = 0;
    ~fdb_matrix() { delete part; }
};

struct fdb_vector {
    fdb::DevBuf<double> d;
    int64_t n = 0;
};

namespace fdb {
// pattern.cu
int build_pattern(fdb_space* s, int symmetric);
int build_forcing_map(fdb_space* s);
int build_transpose_perm(fdb_space* s, Pattern* p);
int ensure_fused_plan(fdb_space* s, Pattern* p);
int node_bounding_box(fdb_space* s, double lo[3], double hi[3]);
// assemble.cu
struct OpCanon;  // canonical operator (see assemble.cu)
int assemble_operator(fdb_space* s, const fdb_opdesc* op, fdb_matrix* A);
int upload_tensor_constants(int M, int R, const double* host, int count, cudaStream_t st);
int assemble_forcing(fdb_space* s, const double* f_quad_dev, double* b_dev);
int quadrature_nodes(fdb_space* s, double* out_dev);
int dofs_coords(fdb_space* s, double* out_dev);
int apply_dirichlet(fdb_matrix* A, const double* g, double* b, double* x0);
// surface.cu: per-cell kernels of manifold spaces (M = 2, N = 3)
int surface_local_assemble(fdb_space* s, const Pattern& P, const OpCanon& op, double* contrib);
int surface_local_forcing(fdb_space* s, const double* f_quad, const int32_t* pos, double* contrib);
int surface_quadrature_nodes(fdb_space* s, double* out);
int surface_dof_coords(fdb_space* s, int first_slot, const int32_t* first, double* out);
int surface_bin_cells(fdb_space* s, const GridDesc& G, int32_t* counter, int32_t* bin_cells, bool fill);
int surface_locate(fdb_space* s, const GridDesc& G, int64_t n_locs, const double* locs_d, int32_t* ids_d);
int surface_eval_pointwise(fdb_space* s, int64_t n_locs, const double* locs_d, const int32_t* ids_d, int32_t* cols, double* vals);
int surface_cell_basis_integrals(fdb_space* s, double* integ, double* meas);
// evaluate.cu (host arrays in, host arrays out)
int locate_host(fdb_space* s, int64_t n_locs, const double* locs, int32_t* ids);
int eval_pointwise_host(fdb_space* s, int64_t n_locs, const double* locs, int32_t* ids, int32_t* cols, double* vals);
int eval_areal_host(fdb_space* s, int n_sub, const double* incidence, int64_t capacity, int64_t* n_triplets, int32_t* rows,
                    int32_t* cols, double* vals, double* D);
// comm.cu
int halo_exchange(fdb_matrix* A, double* vec);
int allreduce_sum(fdb_matrix* A, const double* in, double* out, int count);
// solve.cu
int spmv(fdb_matrix* A, const double* x, double* y);
int solve(fdb_matrix* A, const double* b, double* x, const fdb_solver_opts* opts, fdb_solve_stats* stats);
// topology.cu
int enumerate_dofs(int M, int R, int n_nodes, int n_cells, const int32_t* cells_rowmajor, const uint8_t* boundary_nodes,
                   int32_t* dofs_colmajor, uint8_t* boundary_dofs, int* n_dofs);
}  // namespace fdb
