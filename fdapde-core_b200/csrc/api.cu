// extern "C" surface of libfdapde_b200.so (declared in include/fdapde_b200.h).
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>

#include "local_matrix.cuh"
#include "solve_common.cuh"

namespace fdb {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

// ---- caching device allocator -----------------------------------------------------------------------------------------
struct Pool {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void*> free_blocks;   // (device, size) -> block
    std::unordered_map<void*, std::pair<int, size_t>> live;     // block -> (device, size)
};
static Pool& pool() {
    static Pool* p = new Pool();  // leaked on purpose: must outlive every static DevBuf
    return *p;
}
static size_t round_up(size_t bytes) {
    const size_t g = bytes < (1u << 20) ? 512 : (2u << 20);  // 512 B below 1 MiB, 2 MiB above
    return (bytes + g - 1) / g * g;
}
void* pool_alloc(size_t bytes) {
    const size_t want = round_up(bytes);
    Pool& P = pool();
    int dev = 0;
    {   // blocks are only ever reused on the device they were allocated on; and there is no CPU fallback
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error(std::string("no CUDA device: ") + cudaGetErrorString(e));
            return nullptr;
        }
    }
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.free_blocks.lower_bound({dev, want});
        if (it != P.free_blocks.end() && it->first.first == dev &&
            it->first.second <= want + want / 4 + (1u << 20)) {  // close enough fit
            void* p = it->second;
            P.live[p] = it->first;
            P.free_blocks.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {  // give cached blocks back to the driver and retry once
        cudaGetLastError();
        fdb_trim();
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) {
        set_error(std::string("cudaMalloc(") + std::to_string(want) + " B): " + cudaGetErrorString(e));
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(P.mu);
    P.live[p] = {dev, want};
    return p;
}
void pool_free(void* p) {
    if (!p) return;
    Pool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.live.find(p);
    if (it == P.live.end()) return;
    P.free_blocks.emplace(it->second, p);
    P.live.erase(it);
}

// ---- caching pinned-host allocator -------------------------------------------------------------------------------------
// Result arrays handed to the host (CSC values, pattern) are DMA targets: page-locked memory takes the copy at PCIe
// speed, pageable memory is staged by the driver at a fraction of it.  cudaHostAlloc itself is slow (it pins pages), so
// released blocks are cached like the device blocks above.
struct HostPool {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;
    std::unordered_map<void*, size_t> live;
};
static HostPool& host_pool() {
    static HostPool* p = new HostPool();
    return *p;
}
}  // namespace fdb

using namespace fdb;

extern "C" {

const char* fdb_last_error(void) { return g_last_error.c_str(); }
int fdb_version(void) { return 100; }

int fdb_trim(void) {
    Pool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    cudaDeviceSynchronize();
    for (auto& kv : P.free_blocks) cudaFree(kv.second);
    P.free_blocks.clear();
    HostPool& H = host_pool();
    std::lock_guard<std::mutex> lh(H.mu);
    for (auto& kv : H.free_blocks) cudaFreeHost(kv.second);
    H.free_blocks.clear();
    return FDB_OK;
}

int fdb_host_alloc(size_t bytes, void** out) {
    FDB_CHECK(out, FDB_ERR_ARG, "null argument");
    *out = nullptr;
    const size_t want = bytes < (1u << 20) ? ((bytes + 4095) / 4096 * 4096) : round_up(bytes);
    HostPool& P = host_pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.free_blocks.lower_bound(want);
        if (it != P.free_blocks.end() && it->first <= want + want / 4 + (1u << 20)) {
            *out = it->second;
            P.live[*out] = it->first;
            P.free_blocks.erase(it);
            return FDB_OK;
        }
    }
    void* p = nullptr;
    FDB_CUDA(cudaHostAlloc(&p, want ? want : 4096, cudaHostAllocPortable));
    std::lock_guard<std::mutex> lk(P.mu);
    P.live[p] = want;
    *out = p;
    return FDB_OK;
}
int fdb_host_free(void* p) {
    if (!p) return FDB_OK;
    HostPool& P = host_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.live.find(p);
    FDB_CHECK(it != P.live.end(), FDB_ERR_ARG, "fdb_host_free: not a block of fdb_host_alloc");
    P.free_blocks.emplace(it->second, p);
    P.live.erase(it);
    return FDB_OK;
}

int fdb_device_count(int* count) {
    FDB_CHECK(count, FDB_ERR_ARG, "null argument");
    *count = 0;
    FDB_CUDA(cudaGetDeviceCount(count));
    return FDB_OK;
}
int fdb_set_device(int device) {
    FDB_CUDA(cudaSetDevice(device));
    return FDB_OK;
}

// ---- space --------------------------------------------------------------------------------------------------------
__global__ void k_transpose_cells(int n_cells, int nv, const int32_t* __restrict__ rowmajor, int32_t* __restrict__ soa) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_cells * nv) return;
    int e = (int)(t / nv), k = (int)(t % nv);
    soa[(size_t)k * n_cells + e] = rowmajor[t];
}

__global__ void k_pack_coords(int n_nodes, int N, int pk, const double* __restrict__ soa, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    for (int d = 0; d < pk; ++d) out[(size_t)i * pk + d] = d < N ? soa[(size_t)d * n_nodes + i] : 0.0;
}

int fdb_space_create(fdb_space** out, int M, int N, int R, int n_nodes, int n_cells, const double* nodes,
                     const int32_t* cells, int n_dofs, const int32_t* dofs) {
    FDB_CHECK(out, FDB_ERR_ARG, "null output handle");
    *out = nullptr;
    FDB_CHECK(M == N || (M == 2 && N == 3), FDB_ERR_UNSUPPORTED,
              "supported meshes: Triangulation<2,2>, <3,3> and the surface case <2,3>");
    FDB_CHECK(nodes && dofs, FDB_ERR_ARG, "null mesh arrays");
    FDB_CHECK(n_nodes > 0 && n_cells > 0 && n_dofs >= n_nodes, FDB_ERR_ARG, "bad mesh sizes");
    fdb_space* s = new fdb_space();
    int rc = build_fe_tables(M, R, &s->tab_host, &s->poly_host);
    if (rc != FDB_OK) { delete s; return rc; }
    s->M = M; s->N = N; s->R = R;
    s->nb = s->tab_host.nb;
    s->nq = s->tab_host.nq;
    s->n_nodes = n_nodes; s->n_cells = n_cells; s->n_dofs = n_dofs;
    auto fail = [&](int code) { fdb_space_destroy(s); return code; };
    cudaError_t e = cudaGetDevice(&s->device);
    if (e != cudaSuccess) { set_error(std::string("no CUDA device: ") + cudaGetErrorString(e)); return fail(FDB_ERR_CUDA); }
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
    e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { set_error(std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); return fail(FDB_ERR_CUDA); }
    s->own_stream = true;
#define FDB_SPACE_TRY(x) do { int rc_ = (x); if (rc_ != FDB_OK) return fail(rc_); } while (0)
#define FDB_SPACE_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_error(std::string(#x) + ": " + cudaGetErrorString(e_)); return fail(FDB_ERR_CUDA); } } while (0)
    FDB_SPACE_TRY(s->tab.alloc(1));
    FDB_SPACE_CUDA(cudaMemcpyAsync(s->tab.p, &s->tab_host, sizeof(FeTables), cudaMemcpyHostToDevice, s->stream));
    if (M == N) {   // reference tensors of the constant-coefficient form (local_matrix.cuh: tens_entry), contracted over the
        // quadrature rule once: T^mn_ij = sum_q w_q d_m psi_i d_n psi_j, A^n_ij = sum_q w_q psi_i d_n psi_j, R_ij = sum_q w_q psi_i psi_j
        const FeTables& T = s->tab_host;
        const int NB = T.nb, TS = tens_stride(M), TL = tens_stride_of(M, MODE_TENS_LAP);
        std::vector<double> h((size_t)tens_total(M, NB), 0.0);
        for (int q = 0; q < T.nq; ++q)
            for (int i = 0; i < NB; ++i)
                for (int j = 0; j < NB; ++j) {
                    double* t = h.data() + (size_t)(i * NB + j) * TS;
                    for (int m = 0; m < M; ++m)
                        for (int n = 0; n < M; ++n)
                            t[m * M + n] += T.w[q] * T.gref[(q * NB + i) * M + m] * T.gref[(q * NB + j) * M + n];
                    for (int n = 0; n < M; ++n) t[M * M + n] += T.w[q] * T.phi[q * NB + i] * T.gref[(q * NB + j) * M + n];
                    t[M * M + M] += T.w[q] * T.phi[q * NB + i] * T.phi[q * NB + j];
                }
        // specialised rows: Laplacian only (W symmetric -> T^mn + T^nm for m < n), reaction only (R_ij)
        double* hl = h.data() + tens_offset_of(M, NB, MODE_TENS_LAP);
        double* hr = h.data() + tens_offset_of(M, NB, MODE_TENS_REAC);
        for (int ij = 0; ij < NB * NB; ++ij) {
            const double* t = h.data() + (size_t)ij * TS;
            int k = 0;
            for (int m = 0; m < M; ++m)
                for (int n = m; n < M; ++n) hl[(size_t)ij * TL + k++] = (m == n) ? t[m * M + m] : t[m * M + n] + t[n * M + m];
            hr[ij] = t[M * M + M];
        }
        FDB_SPACE_TRY(s->tens.alloc(h.size()));
        FDB_SPACE_CUDA(cudaMemcpyAsync(s->tens.p, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, s->stream));
        FDB_SPACE_TRY(upload_tensor_constants(M, s->R, h.data(), (int)h.size(), s->stream));
        FDB_SPACE_CUDA(cudaStreamSynchronize(s->stream));   // h is scoped to this block
    }
    FDB_SPACE_TRY(s->poly.alloc(1));
    FDB_SPACE_CUDA(cudaMemcpyAsync(s->poly.p, &s->poly_host, sizeof(PolyTables), cudaMemcpyHostToDevice, s->stream));
    // Eigen's column-major node matrix and dof table ARE struct-of-arrays: upload as they are
    FDB_SPACE_TRY(s->coords.alloc((size_t)n_nodes * N));
    FDB_SPACE_CUDA(cudaMemcpyAsync(s->coords.p, nodes, sizeof(double) * (size_t)n_nodes * N, cudaMemcpyHostToDevice, s->stream));
    FDB_SPACE_TRY(s->dofs.alloc((size_t)n_cells * s->nb));
    FDB_SPACE_CUDA(cudaMemcpyAsync(s->dofs.p, dofs, sizeof(int32_t) * (size_t)n_cells * s->nb, cudaMemcpyHostToDevice, s->stream));
    {   // packed per-node copy for the gathers of the fused kernel
        const int pk = (N == 3) ? 4 : 2;
        FDB_SPACE_TRY(s->coords_pk.alloc((size_t)n_nodes * pk));
        k_pack_coords<<<(unsigned)((n_nodes + 255) / 256), 256, 0, s->stream>>>(n_nodes, N, pk, s->coords.p, s->coords_pk.p);
        FDB_SPACE_CUDA(cudaGetLastError());
        if (const char* e = getenv("FDB_FUSED_THREADS")) s->fused_threads = atoi(e) > 0 ? atoi(e) : 0;
    }
    if (cells) {  // row-major cells -> SoA on the device
        DevBuf<int32_t> tmp;
        FDB_SPACE_TRY(tmp.alloc((size_t)n_cells * (M + 1)));
        FDB_SPACE_TRY(s->verts.alloc((size_t)n_cells * (M + 1)));
        FDB_SPACE_CUDA(cudaMemcpyAsync(tmp.p, cells, sizeof(int32_t) * (size_t)n_cells * (M + 1), cudaMemcpyHostToDevice, s->stream));
        int64_t tot = (int64_t)n_cells * (M + 1);
        k_transpose_cells<<<(unsigned)((tot + 255) / 256), 256, 0, s->stream>>>(n_cells, M + 1, tmp.p, s->verts.p);
        FDB_SPACE_CUDA(cudaGetLastError());
        FDB_SPACE_CUDA(cudaStreamSynchronize(s->stream));
        s->verts_p = s->verts.p;
    } else {
        s->verts_p = s->dofs.p;  // first M+1 dof columns are the vertices (lagrangian_basis.h:96,102-103)
    }
    FDB_SPACE_CUDA(cudaStreamSynchronize(s->stream));
#undef FDB_SPACE_TRY
#undef FDB_SPACE_CUDA
    *out = s;
    return FDB_OK;
}

// A space is shared by the matrices assembled on it (they point into its pattern); it is reference counted so that
// handles may be released in any order (the reference's PDE objects are copyable: pde.h:167-169, type_erasure.h:222).
void fdb_space_destroy(fdb_space* s) {
    if (!s) return;
    if (--s->refs > 0) return;
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    for (int k = 0; k < 3; ++k)
        if (s->ev[k]) cudaEventDestroy(s->ev[k]);
    delete s;
}

int fdb_space_set_stream(fdb_space* s, void* cuda_stream) {
    FDB_CHECK(s, FDB_ERR_ARG, "null space");
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    s->stream = (cudaStream_t)cuda_stream;
    s->own_stream = false;
    return FDB_OK;
}

int fdb_space_sync(fdb_space* s) {
    FDB_CHECK(s, FDB_ERR_ARG, "null space");
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    return FDB_OK;
}

int fdb_space_info(const fdb_space* s, int* n_dofs, int* n_cells, int* n_basis, int* n_quad) {
    FDB_CHECK(s, FDB_ERR_ARG, "null space");
    if (n_dofs) *n_dofs = s->n_dofs;
    if (n_cells) *n_cells = s->n_cells;
    if (n_basis) *n_basis = s->nb;
    if (n_quad) *n_quad = s->nq;
    return FDB_OK;
}

int fdb_space_set_fused(fdb_space* s, int enabled) {
    FDB_CHECK(s, FDB_ERR_ARG, "null space");
    s->force_two_kernel = !enabled;
    return FDB_OK;
}

int fdb_space_set_profiling(fdb_space* s, int enabled) {
    FDB_CHECK(s, FDB_ERR_ARG, "null space");
    if (enabled && !s->ev[0])
        for (int k = 0; k < 3; ++k) FDB_CUDA(cudaEventCreate(&s->ev[k]));
    s->profile = enabled != 0;
    s->ev_valid = false;
    return FDB_OK;
}

int fdb_space_last_timings(fdb_space* s, double* ms, int capacity, int* count) {
    FDB_CHECK(s && ms && count, FDB_ERR_ARG, "null argument");
    FDB_CHECK(s->profile && s->ev_valid, FDB_ERR_STATE, "no profiled assembly has run");
    FDB_CUDA(cudaEventSynchronize(s->ev[2]));
    *count = 0;
    for (int k = 0; k < 2 && k < capacity; ++k) {
        float t = 0;
        FDB_CUDA(cudaEventElapsedTime(&t, s->ev[k], s->ev[k + 1]));
        ms[k] = t;
        ++*count;
    }
    return FDB_OK;
}

int fdb_space_last_path(const fdb_space* s, int* fused, int* launches) {
    FDB_CHECK(s, FDB_ERR_ARG, "null space");
    FDB_CHECK(s->last_fused >= 0, FDB_ERR_STATE, "no assembly has run on this space");
    if (fused) *fused = s->last_fused;
    if (launches) *launches = s->last_launches;
    return FDB_OK;
}

int fdb_space_set_boundary(fdb_space* s, const uint8_t* boundary_dofs) {
    FDB_CHECK(s && boundary_dofs, FDB_ERR_ARG, "null argument");
    if (!s->boundary.p) FDB_TRY(s->boundary.alloc(s->n_dofs));
    FDB_CUDA(cudaMemcpyAsync(s->boundary.p, boundary_dofs, s->n_dofs, cudaMemcpyHostToDevice, s->stream));
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    s->has_boundary = true;
    return FDB_OK;
}

int fdb_space_set_dof0_rule(fdb_space* s, int enabled) {
    FDB_CHECK(s, FDB_ERR_ARG, "null space");
    s->dof0_rule = enabled != 0;
    return FDB_OK;
}

int fdb_enumerate_dofs(int M, int R, int n_nodes, int n_cells, const int32_t* cells, const uint8_t* boundary_nodes,
                       int32_t* dofs, uint8_t* boundary_dofs, int* n_dofs) {
    return enumerate_dofs(M, R, n_nodes, n_cells, cells, boundary_nodes, dofs, boundary_dofs, n_dofs);
}

int fdb_quadrature_nodes(fdb_space* s, double* out) {
    FDB_CHECK(s && out, FDB_ERR_ARG, "null argument");
    size_t count = (size_t)s->n_cells * s->nq * s->N;
    DevBuf<double> d;
    FDB_TRY(d.alloc(count));
    FDB_TRY(quadrature_nodes(s, d.p));
    FDB_CUDA(cudaMemcpyAsync(out, d.p, sizeof(double) * count, cudaMemcpyDeviceToHost, s->stream));
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    return FDB_OK;
}

int fdb_dofs_coords(fdb_space* s, double* out) {
    FDB_CHECK(s && out, FDB_ERR_ARG, "null argument");
    size_t count = (size_t)s->n_dofs * s->N;
    DevBuf<double> d;
    FDB_TRY(d.alloc(count));
    FDB_TRY(dofs_coords(s, d.p));
    FDB_CUDA(cudaMemcpyAsync(out, d.p, sizeof(double) * count, cudaMemcpyDeviceToHost, s->stream));
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    return FDB_OK;
}

// ---- point location and basis evaluation (next-row N1) ------------------------------------------------------------
int fdb_locate(fdb_space* s, int64_t n_locs, const double* locs, int32_t* cell_ids) {
    return locate_host(s, n_locs, locs, cell_ids);
}
int fdb_eval_pointwise(fdb_space* s, int64_t n_locs, const double* locs, int32_t* cell_ids, int32_t* cols, double* vals) {
    return eval_pointwise_host(s, n_locs, locs, cell_ids, cols, vals);
}
int fdb_eval_areal(fdb_space* s, int n_subdomains, const double* incidence, int64_t capacity, int64_t* n_triplets,
                   int32_t* rows, int32_t* cols, double* vals, double* measures) {
    return eval_areal_host(s, n_subdomains, incidence, capacity, n_triplets, rows, cols, vals, measures);
}

// ---- pattern ------------------------------------------------------------------------------------------------------
int fdb_pattern_nnz(fdb_space* s, int symmetric, int64_t* nnz) {
    FDB_CHECK(s && nnz, FDB_ERR_ARG, "null argument");
    FDB_TRY(build_pattern(s, symmetric ? 1 : 0));
    *nnz = s->pat[symmetric ? 1 : 0].nnz;
    return FDB_OK;
}

int fdb_space_prepare(fdb_space* s, int symmetric) {
    FDB_CHECK(s, FDB_ERR_ARG, "null space");
    FDB_TRY(build_pattern(s, symmetric ? 1 : 0));
    if (!s->force_two_kernel) FDB_TRY(ensure_fused_plan(s, &s->pat[symmetric ? 1 : 0]));
    return FDB_OK;
}

int fdb_pattern_download(fdb_space* s, int symmetric, int32_t* outer, int32_t* inner) {
    FDB_CHECK(s && outer && inner, FDB_ERR_ARG, "null argument");
    FDB_TRY(build_pattern(s, symmetric ? 1 : 0));
    const Pattern& P = s->pat[symmetric ? 1 : 0];
    // structurally symmetric pattern: CSR(rowptr, colidx) == CSC(outer, inner)
    FDB_CUDA(cudaMemcpyAsync(outer, P.rowptr.p, sizeof(int32_t) * ((size_t)s->n_dofs + 1), cudaMemcpyDeviceToHost, s->stream));
    FDB_CUDA(cudaMemcpyAsync(inner, P.colidx.p, sizeof(int32_t) * (size_t)P.nnz, cudaMemcpyDeviceToHost, s->stream));
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    return FDB_OK;
}

// ---- matrix -------------------------------------------------------------------------------------------------------
int fdb_matrix_create(fdb_space* s, fdb_matrix** out) {
    FDB_CHECK(s && out, FDB_ERR_ARG, "null argument");
    fdb_matrix* A = new fdb_matrix();
    A->space = s;
    ++s->refs;
    *out = A;
    return FDB_OK;
}
void fdb_matrix_destroy(fdb_matrix* A) {
    if (!A) return;
    fdb_space* s = A->space;
    if (s && s->stream) cudaStreamSynchronize(s->stream);
    delete A;
    fdb_space_destroy(s);  // drops the matrix's reference
}
int fdb_matrix_nnz(const fdb_matrix* A, int64_t* nnz) {
    FDB_CHECK(A && nnz, FDB_ERR_ARG, "null argument");
    FDB_CHECK(A->pat, FDB_ERR_STATE, "matrix has not been assembled");
    *nnz = A->pat->nnz;
    return FDB_OK;
}

int fdb_assemble_operator(fdb_space* s, const fdb_opdesc* op, fdb_matrix* A) { return assemble_operator(s, op, A); }

__global__ void k_gather(int64_t n, const int32_t* __restrict__ perm, const double* __restrict__ in, double* __restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = in[perm[t]];
}

int fdb_matrix_download_csc(fdb_matrix* A, int32_t* outer, int32_t* inner, double* values) {
    FDB_CHECK(A && A->assembled, FDB_ERR_STATE, "matrix has not been assembled");
    fdb_space* s = A->space;
    Pattern* P = const_cast<Pattern*>(A->pat);
    if (outer) FDB_CUDA(cudaMemcpyAsync(outer, P->rowptr.p, sizeof(int32_t) * ((size_t)s->n_dofs + 1), cudaMemcpyDeviceToHost, s->stream));
    if (inner) FDB_CUDA(cudaMemcpyAsync(inner, P->colidx.p, sizeof(int32_t) * (size_t)P->nnz, cudaMemcpyDeviceToHost, s->stream));
    if (values) {
        // The device holds CSR(A).  CSC(A) has the same index arrays (structurally symmetric pattern) and the values
        // permuted by the transpose map.  This is done for symmetric operators too: Dirichlet rows break symmetry.
        FDB_TRY(build_transpose_perm(s, P));
        DevBuf<double> tmp;
        FDB_TRY(tmp.alloc((size_t)P->nnz));
        k_gather<<<(unsigned)((P->nnz + 255) / 256), 256, 0, s->stream>>>(P->nnz, P->tperm.p, A->val.p, tmp.p);
        FDB_CUDA(cudaGetLastError());
        FDB_CUDA(cudaMemcpyAsync(values, tmp.p, sizeof(double) * (size_t)P->nnz, cudaMemcpyDeviceToHost, s->stream));
        FDB_CUDA(cudaStreamSynchronize(s->stream));
    }
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    return FDB_OK;
}

int fdb_discretize_operator(fdb_space* s, const fdb_opdesc* op, int32_t* outer, int32_t* inner, double* values) {
    FDB_CHECK(s && op, FDB_ERR_ARG, "null argument");
    const int sym = op->symmetric ? 1 : 0;
    FDB_TRY(build_pattern(s, sym));
    const Pattern& P = s->pat[sym];
    // The index arrays are final before any value exists (build_pattern has synchronised its stream): their download
    // runs on a second stream underneath the assembly kernels.
    cudaStream_t cs = nullptr;
    if (outer || inner) {
        FDB_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaError_t e = cudaSuccess;
        if (outer) e = cudaMemcpyAsync(outer, P.rowptr.p, sizeof(int32_t) * ((size_t)s->n_dofs + 1), cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess && inner) e = cudaMemcpyAsync(inner, P.colidx.p, sizeof(int32_t) * (size_t)P.nnz, cudaMemcpyDeviceToHost, cs);
        if (e != cudaSuccess) {
            cudaStreamDestroy(cs);
            set_error(std::string("pattern download: ") + cudaGetErrorString(e));
            return FDB_ERR_CUDA;
        }
    }
    fdb_matrix* A = nullptr;
    int rc = fdb_matrix_create(s, &A);
    if (rc == FDB_OK) rc = assemble_operator(s, op, A);
    if (rc == FDB_OK && values) {
        if (P.symmetric) {
            // a freshly assembled symmetric operator holds the same bits in (r, c) and (c, r) (one sum, stored twice):
            // CSC values == CSR values, no transpose pass
            cudaError_t e = cudaMemcpyAsync(values, A->val.p, sizeof(double) * (size_t)P.nnz, cudaMemcpyDeviceToHost, s->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
            if (e != cudaSuccess) { set_error(std::string("value download: ") + cudaGetErrorString(e)); rc = FDB_ERR_CUDA; }
        } else {
            rc = fdb_matrix_download_csc(A, nullptr, nullptr, values);
        }
    }
    if (cs) {
        cudaError_t e = cudaStreamSynchronize(cs);
        cudaStreamDestroy(cs);
        if (rc == FDB_OK && e != cudaSuccess) { set_error(std::string("pattern download: ") + cudaGetErrorString(e)); rc = FDB_ERR_CUDA; }
    }
    if (A) fdb_matrix_destroy(A);
    return rc;
}

// ---- vectors ------------------------------------------------------------------------------------------------------
int fdb_vector_create(int64_t n, fdb_vector** out) {
    FDB_CHECK(out && n >= 0, FDB_ERR_ARG, "bad argument");
    fdb_vector* v = new fdb_vector();
    int rc = v->d.alloc((size_t)n);
    if (rc != FDB_OK) { delete v; return rc; }
    v->n = n;
    *out = v;
    return FDB_OK;
}
void fdb_vector_destroy(fdb_vector* v) {
    if (!v) return;
    cudaDeviceSynchronize();
    delete v;
}
int fdb_vector_upload(fdb_vector* v, const double* host, int64_t n) {
    FDB_CHECK(v && host && n <= v->n, FDB_ERR_ARG, "bad argument");
    // Vectors carry no stream of their own and every space stream is non-blocking (no implicit ordering with the
    // legacy stream): order the copy against all queued work on both sides, so that a later kernel on a space stream
    // sees the data and an earlier one is not overwritten under its feet.
    FDB_CUDA(cudaDeviceSynchronize());
    FDB_CUDA(cudaMemcpy(v->d.p, host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    FDB_CUDA(cudaDeviceSynchronize());
    return FDB_OK;
}
int fdb_vector_download(const fdb_vector* v, double* host, int64_t n) {
    FDB_CHECK(v && host && n <= v->n, FDB_ERR_ARG, "bad argument");
    FDB_CUDA(cudaDeviceSynchronize());
    FDB_CUDA(cudaMemcpy(host, v->d.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    return FDB_OK;
}
int fdb_vector_fill(fdb_vector* v, double value) {
    FDB_CHECK(v, FDB_ERR_ARG, "null vector");
    FDB_CUDA(cudaDeviceSynchronize());   // same ordering rule as fdb_vector_upload (cudaMemset is asynchronous)
    if (value == 0.0) {
        FDB_CUDA(cudaMemset(v->d.p, 0, sizeof(double) * (size_t)v->n));
    } else {
        std::vector<double> h((size_t)v->n, value);
        FDB_CUDA(cudaMemcpy(v->d.p, h.data(), sizeof(double) * (size_t)v->n, cudaMemcpyHostToDevice));
    }
    FDB_CUDA(cudaDeviceSynchronize());
    return FDB_OK;
}

// ---- forcing ------------------------------------------------------------------------------------------------------
int fdb_assemble_forcing(fdb_space* s, const fdb_vector* f_quad, fdb_vector* b) {
    FDB_CHECK(s && f_quad && b, FDB_ERR_ARG, "null argument");
    FDB_CHECK(f_quad->n >= (int64_t)s->n_cells * s->nq, FDB_ERR_ARG, "forcing needs n_cells * n_quad values");
    FDB_CHECK(b->n >= s->n_dofs, FDB_ERR_ARG, "load vector needs n_dofs entries");
    return assemble_forcing(s, f_quad->d.p, b->d.p);
}

int fdb_discretize_forcing(fdb_space* s, const double* f_host, double* b_host) {
    FDB_CHECK(s && f_host && b_host, FDB_ERR_ARG, "null argument");
    size_t nf = (size_t)s->n_cells * s->nq;
    DevBuf<double> f, b;
    FDB_TRY(f.alloc(nf));
    FDB_TRY(b.alloc(s->n_dofs));
    FDB_CUDA(cudaMemcpyAsync(f.p, f_host, sizeof(double) * nf, cudaMemcpyHostToDevice, s->stream));
    FDB_TRY(assemble_forcing(s, f.p, b.p));
    FDB_CUDA(cudaMemcpyAsync(b_host, b.p, sizeof(double) * s->n_dofs, cudaMemcpyDeviceToHost, s->stream));
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    return FDB_OK;
}

// ---- Dirichlet / solve --------------------------------------------------------------------------------------------
int fdb_set_dirichlet(fdb_matrix* A, const fdb_vector* g, fdb_vector* b, fdb_vector* x0) {
    FDB_CHECK(A && g && b, FDB_ERR_ARG, "null argument");
    FDB_CHECK(g->n >= A->space->n_dofs && b->n >= A->space->n_dofs && (!x0 || x0->n >= A->space->n_dofs), FDB_ERR_ARG,
              "vectors shorter than n_dofs");
    return apply_dirichlet(A, g->d.p, b->d.p, x0 ? x0->d.p : nullptr);
}

__global__ void k_axpby(int64_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                        double* __restrict__ z) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) z[i] = a * x[i] + b * y[i];
}
__global__ void k_set_boundary_values(int n, int dof0_rule, const uint8_t* __restrict__ boundary,
                                      const double* __restrict__ g, double* __restrict__ rhs, double* __restrict__ x) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if ((i == 0 && dof0_rule) || boundary[i]) { rhs[i] = g[i]; x[i] = g[i]; }
}

int fdb_matrix_axpby(fdb_matrix* C, double a, const fdb_matrix* A, double b, const fdb_matrix* B) {
    FDB_CHECK(C && A && B && A->assembled && B->assembled, FDB_ERR_STATE, "matrices must be assembled");
    FDB_CHECK(A->space == B->space && C->space == A->space, FDB_ERR_ARG, "matrices of different spaces");
    FDB_CHECK(A->pat->nnz == B->pat->nnz, FDB_ERR_ARG, "matrices with different patterns");
    const int64_t nnz = A->pat->nnz;
    if (C != A && C != B && (C->pat != A->pat || C->val.n < (size_t)nnz)) {
        FDB_TRY(C->val.alloc((size_t)nnz));
        C->pat = A->pat;
    }
    k_axpby<<<(unsigned)((nnz + 255) / 256), 256, 0, A->space->stream>>>(nnz, a, A->val.p, b, B->val.p, C->val.p);
    FDB_CUDA(cudaGetLastError());
    C->assembled = true;
    ++C->val_version;
    return FDB_OK;
}

int fdb_solve_parabolic(fdb_matrix* stiff, fdb_matrix* mass, double dt, int m, const double* f_quad, const double* g,
                        const double* u0, double* solution, const fdb_solver_opts* opts, fdb_solve_stats* stats) {
    FDB_CHECK(stiff && mass && stiff->assembled && mass->assembled, FDB_ERR_STATE, "solver must be initialized first!");
    FDB_CHECK(stiff->space == mass->space, FDB_ERR_ARG, "stiff and mass belong to different spaces");
    FDB_CHECK(f_quad && u0 && solution && opts && m >= 1 && dt > 0, FDB_ERR_ARG, "bad argument");
    fdb_space* s = stiff->space;
    FDB_CHECK(!g || s->has_boundary, FDB_ERR_STATE, "fdb_space_set_boundary has not been called");
    cudaStream_t st = s->stream;
    const int n = s->n_dofs;
    const size_t nf = (size_t)s->n_cells * s->nq;
    fdb_matrix K, Mdt;
    K.space = Mdt.space = s;
    FDB_TRY(fdb_matrix_axpby(&Mdt, 1.0 / dt, mass, 0.0, mass));        // mass / dt
    FDB_TRY(fdb_matrix_axpby(&K, 1.0, &Mdt, 1.0, stiff));              // K = mass / dt + stiff
    DevBuf<double> u, rhs, force, fq, gd;
    FDB_TRY(u.alloc(n)); FDB_TRY(rhs.alloc(n)); FDB_TRY(force.alloc(n)); FDB_TRY(fq.alloc(nf));
    if (g) FDB_TRY(gd.alloc(n));
    FDB_CUDA(cudaMemcpyAsync(u.p, u0, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    memcpy(solution, u0, sizeof(double) * n);
    if (g) {  // Dirichlet rows of K (values only, pattern kept): any column of g works, only the matrix matters here
        FDB_CUDA(cudaMemcpyAsync(gd.p, g, sizeof(double) * n, cudaMemcpyHostToDevice, st));
        FDB_TRY(apply_dirichlet(&K, gd.p, rhs.p, nullptr));
    }
    fdb_solve_stats total{0, 1, 0.0, 0.0};
    // A step that misses the tolerance does not stop the time loop: like the reference (factor once, solve every step,
    // fem_linear_parabolic_solver.h:55-71) every column of the solution is written; the call then returns
    // FDB_ERR_NOT_CONVERGED with stats->converged = 0.
    int rc = FDB_OK;
    bool missed = false;
    for (int i = 0; i + 1 < m && rc == FDB_OK; ++i) {
        FDB_CUDA(cudaMemcpyAsync(fq.p, f_quad + nf * (size_t)(i + 1), sizeof(double) * nf, cudaMemcpyHostToDevice, st));
        FDB_TRY(assemble_forcing(s, fq.p, force.p));                   // force_{i+1}
        FDB_TRY(spmv(&Mdt, u.p, rhs.p));                               // (mass / dt) u_i
        k_axpby<<<(n + 255) / 256, 256, 0, st>>>(n, 1.0, rhs.p, 1.0, force.p, rhs.p);
        if (g) {
            FDB_CUDA(cudaMemcpyAsync(gd.p, g + (size_t)n * (i + 1), sizeof(double) * n, cudaMemcpyHostToDevice, st));
            k_set_boundary_values<<<(n + 255) / 256, 256, 0, st>>>(n, s->dof0_rule ? 1 : 0, s->boundary.p, gd.p, rhs.p, u.p);
        }
        FDB_CUDA(cudaGetLastError());
        fdb_solve_stats one{};
        rc = solve(&K, rhs.p, u.p, opts, &one);                        // warm start from u_i (boundary rows = g)
        total.iters += one.iters;
        total.seconds += one.seconds;
        total.rel_resid = one.rel_resid > total.rel_resid ? one.rel_resid : total.rel_resid;
        if (rc == FDB_ERR_NOT_CONVERGED) { total.converged = 0; missed = true; rc = FDB_OK; }
        FDB_CUDA(cudaMemcpyAsync(solution + (size_t)n * (i + 1), u.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    }
    FDB_CUDA(cudaStreamSynchronize(st));
    K.space = Mdt.space = nullptr;  // stack objects: nothing to release through the handle API
    if (stats) *stats = total;
    if (rc == FDB_OK && missed) {
        set_error("parabolic solve: at least one time step did not reach the tolerance (all columns are filled)");
        rc = FDB_ERR_NOT_CONVERGED;
    }
    return rc;
}

int fdb_set_persistent_cg(int mode) {
    FDB_CHECK(mode >= 0 && mode <= 2, FDB_ERR_ARG, "mode must be 0 (never), 1 (multi-GPU only) or 2 (always)");
    persistent_mode() = mode;
    return FDB_OK;
}

int fdb_set_persistent_sell(int mode) {
    FDB_CHECK(mode >= 0 && mode <= 2, FDB_ERR_ARG, "mode must be 0 (never), 1 (partitioned matrices with a peer plan) or 2 (always)");
    peer_sell_mode() = mode;
    return FDB_OK;
}

int fdb_spmv(fdb_matrix* A, const fdb_vector* x, fdb_vector* y) {
    FDB_CHECK(A && x && y, FDB_ERR_ARG, "null argument");
    FDB_CHECK(x->n >= A->space->n_dofs && y->n >= A->space->n_dofs, FDB_ERR_ARG, "vectors shorter than n_dofs");
    return spmv(A, x->d.p, y->d.p);
}

int fdb_solve(fdb_matrix* A, const fdb_vector* b, fdb_vector* x, const fdb_solver_opts* opts, fdb_solve_stats* stats) {
    FDB_CHECK(A && b && x, FDB_ERR_ARG, "null argument");
    FDB_CHECK(A->space && b->n >= A->space->n_dofs && x->n >= A->space->n_dofs, FDB_ERR_ARG, "vectors shorter than n_dofs");
    return solve(A, b->d.p, x->d.p, opts, stats);
}

int fdb_solve_host(fdb_matrix* A, const double* b_host, double* x_host, const fdb_solver_opts* opts,
                   fdb_solve_stats* stats) {
    FDB_CHECK(A && b_host && x_host, FDB_ERR_ARG, "null argument");
    fdb_space* s = A->space;
    DevBuf<double> b, x;
    FDB_TRY(b.alloc(s->n_dofs));
    FDB_TRY(x.alloc(s->n_dofs));
    FDB_CUDA(cudaMemcpyAsync(b.p, b_host, sizeof(double) * s->n_dofs, cudaMemcpyHostToDevice, s->stream));
    FDB_CUDA(cudaMemcpyAsync(x.p, x_host, sizeof(double) * s->n_dofs, cudaMemcpyHostToDevice, s->stream));
    int rc = solve(A, b.p, x.p, opts, stats);
    if (rc == FDB_OK || rc == FDB_ERR_NOT_CONVERGED) {
        cudaMemcpyAsync(x_host, x.p, sizeof(double) * s->n_dofs, cudaMemcpyDeviceToHost, s->stream);
        cudaStreamSynchronize(s->stream);
    }
    return rc;
}

}  // extern "C"
