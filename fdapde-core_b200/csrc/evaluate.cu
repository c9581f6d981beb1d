// Next-row N1: point location and the basis-evaluation matrices Psi.
//
// Replaces the reference's serial loops
//   Triangulation::locate -> TreeSearch::locate      geometry/triangulation.h:252-255, tree_search.h:28-86
//   Simplex::contains                                geometry/simplex.h:115-128
//   pointwise_evaluation<LagrangianBasis>::eval      basis/lagrangian_basis.h:203-235
//   areal_evaluation<LagrangianBasis>::eval          basis/lagrangian_basis.h:238-283
//   Integrator::integrate_cell                       utils/integration/integrator.h:45-59
// Point location: the reference moves every cell's bounding box to a point of R^{2N} and range-searches a KD-tree
// (an alternating digital tree) once per query.  Here the boxes are binned into a uniform grid built by count / scan /
// fill (two passes over the cells, no tree, no pointers), and one thread per query point tests the cells of its bin
// with the reference's own barycentric criterion.  A point shared by several cells (on an edge / vertex) gets the
// SMALLEST containing cell id -- the reference's winner depends on std::unordered_set iteration order.
// Psi rows are embarrassingly parallel: one thread per point (pointwise) or per cell (areal); results leave as the
// reference's triplet list in emission order, so the host-side setFromTriplets reproduces its matrix entry for entry.
#include <cub/cub.cuh>

#include <algorithm>
#include <climits>
#include <cmath>
#include <vector>

#include "local_matrix.cuh"

namespace fdb {

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// bins overlapped by the (slightly inflated) bounding box of every cell: FILL == false counts, FILL == true writes
template <int M, bool FILL>
__global__ void k_bin_cells(int n_cells, int n_nodes, const int32_t* __restrict__ verts, const double* __restrict__ coords,
                            GridDesc G, int32_t* __restrict__ counter, int32_t* __restrict__ bin_cells) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    double x[M + 1][M];
    gather_vertices<M>(e, n_cells, n_nodes, verts, coords, x);
    int b0[3] = {0, 0, 0}, b1[3] = {0, 0, 0};
    for (int d = 0; d < M; ++d) {
        double lo = x[0][d], hi = x[0][d];
        for (int k = 1; k <= M; ++k) { lo = fmin(lo, x[k][d]); hi = fmax(hi, x[k][d]); }
        b0[d] = bin_of(G, d, lo - G.eps[d]);
        b1[d] = bin_of(G, d, hi + G.eps[d]);
    }
    for (int c = b0[2]; c <= b1[2]; ++c)
        for (int b = b0[1]; b <= b1[1]; ++b)
            for (int a = b0[0]; a <= b1[0]; ++a) {
                const int bin = (c * G.g[1] + b) * G.g[0] + a;
                const int slot = atomicAdd(&counter[bin], 1);
                if (FILL) bin_cells[slot] = e;  // order inside a bin is irrelevant: queries take the minimum id
            }
}

// Simplex::contains: z = (1 - sum, invJ (x - v0)); outside iff some z < -10 eps (simplex.h:121-123, symbols.h:164)
template <int M>
__device__ __forceinline__ bool cell_contains(const Geo<M>& g, const double* p) {
    const double meps = 10 * 2.220446049250313e-16;
    double sum = 0;
    bool in = true;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double t = 0;
#pragma unroll
        for (int r = 0; r < M; ++r) t += g.invJ[m][r] * (p[r] - g.x0[r]);
        sum += t;
        in = in && !(t < -meps);
    }
    return in && !((1 - sum) < -meps);
}

template <int M>
__global__ void k_locate(int64_t n_locs, const double* __restrict__ locs, int n_cells, int n_nodes,
                         const int32_t* __restrict__ verts, const double* __restrict__ coords, GridDesc G,
                         const int32_t* __restrict__ bin_ptr, const int32_t* __restrict__ bin_cells,
                         int32_t* __restrict__ ids) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_locs) return;
    double p[M];
    int bin = 0;
    for (int d = M - 1; d >= 0; --d) {
        p[d] = locs[(size_t)d * n_locs + i];
        bin = bin * G.g[d] + bin_of(G, d, p[d]);
    }
    int best = INT_MAX;
    for (int t = bin_ptr[bin]; t < bin_ptr[bin + 1]; ++t) {
        const int e = bin_cells[t];
        if (e >= best) continue;
        Geo<M> g;
        load_geometry<M>(e, n_cells, n_nodes, verts, coords, g);
        if (cell_contains<M>(g, p)) best = e;
    }
    ids[i] = best == INT_MAX ? -1 : best;
}

// one thread per point: xi = invJ (p - v0), then the n_basis values and their dof columns
template <int M>
__global__ void k_eval_pointwise(int64_t n_locs, const double* __restrict__ locs, const int32_t* __restrict__ ids,
                                 int n_cells, int n_nodes, const int32_t* __restrict__ verts,
                                 const double* __restrict__ coords, const int32_t* __restrict__ dofs,
                                 const PolyTables* __restrict__ poly, int32_t* __restrict__ cols,
                                 double* __restrict__ vals) {
    __shared__ PolyTables P;
    for (int k = threadIdx.x; k < (int)(sizeof(PolyTables) / sizeof(int)); k += blockDim.x)
        reinterpret_cast<int*>(&P)[k] = reinterpret_cast<const int*>(poly)[k];
    __syncthreads();
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_locs) return;
    const int nb = P.nb;
    const int e = ids[i];
    if (e < 0) {
        for (int h = 0; h < nb; ++h) { cols[i * nb + h] = -1; vals[i * nb + h] = 0.0; }
        return;
    }
    Geo<M> g;
    load_geometry<M>(e, n_cells, n_nodes, verts, coords, g);
    double xi[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        double t = 0;
#pragma unroll
        for (int r = 0; r < M; ++r) t += g.invJ[m][r] * (locs[(size_t)r * n_locs + i] - g.x0[r]);
        xi[m] = t;
    }
    for (int h = 0; h < nb; ++h) {
        cols[i * nb + h] = dofs[(size_t)h * n_cells + e];
        vals[i * nb + h] = poly_eval(P, h, xi);
    }
}

// one thread per cell: integrals of the n_basis functions over the cell (integrate_cell: quadrature nodes mapped to the
// cell and back through invJ, exactly as the reference's lambda does) and the cell measure
template <int M>
__global__ void k_cell_basis_integrals(int n_cells, int n_nodes, const int32_t* __restrict__ verts,
                                       const double* __restrict__ coords, const FeTables* __restrict__ tab,
                                       const PolyTables* __restrict__ poly, double* __restrict__ integ,
                                       double* __restrict__ meas) {
    __shared__ PolyTables P;
    __shared__ FeTables T;
    for (int k = threadIdx.x; k < (int)(sizeof(PolyTables) / sizeof(int)); k += blockDim.x)
        reinterpret_cast<int*>(&P)[k] = reinterpret_cast<const int*>(poly)[k];
    stage_tables(tab, &T);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    Geo<M> g;
    load_geometry<M>(e, n_cells, n_nodes, verts, coords, g);
    const int nb = P.nb;
    for (int h = 0; h < nb; ++h) {
        double value = 0;
        for (int q = 0; q < T.nq; ++q) {
            double p[M], xi[M];
#pragma unroll
            for (int r = 0; r < M; ++r) {
                double t = 0;
#pragma unroll
                for (int m = 0; m < M; ++m) t += g.J[r][m] * T.qn[q * M + m];
                p[r] = t + g.x0[r];
            }
#pragma unroll
            for (int m = 0; m < M; ++m) {
                double t = 0;
#pragma unroll
                for (int r = 0; r < M; ++r) t += g.invJ[m][r] * (p[r] - g.x0[r]);
                xi[m] = t;
            }
            value += poly_eval(P, h, xi) * T.w[q];
        }
        integ[(size_t)e * nb + h] = value * g.measure;
    }
    meas[e] = g.measure;
}

// flag[k * n_cells + l] = incidence(k, l) == 1 (k-major: the scan of the flags is the reference's emission order)
__global__ void k_incidence_flags(int n_sub, int n_cells, const double* __restrict__ inc, int32_t* __restrict__ flag) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_sub * n_cells) return;
    int k = (int)(t / n_cells), l = (int)(t % n_cells);
    flag[t] = inc[(size_t)l * n_sub + k] == 1.0 ? 1 : 0;
}

// D_k = measures of the subdomain's cells, added in ascending cell order like the reference's `Di += e.measure()`
__global__ void k_subdomain_measure(int n_sub, int n_cells, const double* __restrict__ inc, const double* __restrict__ meas,
                                    double* __restrict__ D) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_sub) return;
    double s = 0;
    for (int l = 0; l < n_cells; ++l)
        if (inc[(size_t)l * n_sub + k] == 1.0) s += meas[l];
    D[k] = s;
}

__global__ void k_areal_fill(int n_sub, int n_cells, int nb, const int32_t* __restrict__ flag,
                             const int32_t* __restrict__ pos, const int32_t* __restrict__ dofs,
                             const double* __restrict__ integ, const double* __restrict__ D, int32_t* __restrict__ rows,
                             int32_t* __restrict__ cols, double* __restrict__ vals) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_sub * n_cells || !flag[t]) return;
    const int k = (int)(t / n_cells), l = (int)(t % n_cells);
    const int64_t o = (int64_t)pos[t] * nb;
    for (int h = 0; h < nb; ++h) {
        rows[o + h] = k;
        cols[o + h] = dofs[(size_t)h * n_cells + l];
        vals[o + h] = integ[(size_t)l * nb + h] / D[k];
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
static int exclusive_scan(int32_t* d, int64_t n, cudaStream_t st) {
    size_t tb = 0;
    FDB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, d, d, (int)n, st));
    DevBuf<char> tmp;
    FDB_TRY(tmp.alloc(tb));
    FDB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, d, d, (int)n, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    return FDB_OK;
}

static int build_locator(fdb_space* s) {
    Locator& L = s->locator;
    if (L.built) return FDB_OK;
    const bool surface = s->M == 2 && s->N == 3;
    FDB_CHECK(s->M == s->N || surface, FDB_ERR_UNSUPPORTED, "point location: unsupported mesh kind");
    cudaStream_t st = s->stream;
    const int N = s->N;
    // bounding box of the nodes (TriangulationBase::range, triangulation.h:52-56)
    double lo[3], hi[3];
    FDB_TRY(node_bounding_box(s, lo, hi));
    // about two cells per bin, the same number of bins along every axis
    // (a surface fills a 2D sheet of the 3D grid: size the bins from the square root)
    int g = (int)std::ceil(std::pow(std::max(1.0, s->n_cells / 2.0), 1.0 / (surface ? 2 : N)));
    const int gmax = N == 2 ? 4096 : (surface ? 128 : 256);
    g = std::max(1, std::min(g, gmax));
    GridDesc G;
    memset(&G, 0, sizeof(G));
    int64_t n_bins = 1;
    for (int d = 0; d < 3; ++d) {
        G.g[d] = d < N ? g : 1;
        const double ext = d < N ? hi[d] - lo[d] : 1.0;
        G.lo[d] = d < N ? lo[d] : 0.0;
        G.inv_h[d] = (d < N && ext > 0) ? g / ext : 0.0;
        G.eps[d] = 1e-12 * (ext > 0 ? ext : 1.0);   // far above the barycentric tolerance, far below a bin
        n_bins *= G.g[d];
    }
    FDB_TRY(L.bin_ptr.alloc((size_t)n_bins + 1));
    FDB_CUDA(cudaMemsetAsync(L.bin_ptr.p, 0, sizeof(int32_t) * (n_bins + 1), st));
    const int B = 128;
    if (surface) FDB_TRY(surface_bin_cells(s, G, L.bin_ptr.p, nullptr, false));
    else if (s->M == 2) k_bin_cells<2, false><<<grid_for(s->n_cells, B), B, 0, st>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, G, L.bin_ptr.p, nullptr);
    else k_bin_cells<3, false><<<grid_for(s->n_cells, B), B, 0, st>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, G, L.bin_ptr.p, nullptr);
    FDB_CUDA(cudaGetLastError());
    FDB_TRY(exclusive_scan(L.bin_ptr.p, n_bins + 1, st));
    int32_t total = 0;
    FDB_CUDA(cudaMemcpyAsync(&total, L.bin_ptr.p + n_bins, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    FDB_TRY(L.bin_cells.alloc((size_t)std::max(total, 1)));
    DevBuf<int32_t> cursor;
    FDB_TRY(cursor.alloc((size_t)n_bins + 1));
    FDB_CUDA(cudaMemcpyAsync(cursor.p, L.bin_ptr.p, sizeof(int32_t) * (n_bins + 1), cudaMemcpyDeviceToDevice, st));
    if (surface) FDB_TRY(surface_bin_cells(s, G, cursor.p, L.bin_cells.p, true));
    else if (s->M == 2) k_bin_cells<2, true><<<grid_for(s->n_cells, B), B, 0, st>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, G, cursor.p, L.bin_cells.p);
    else k_bin_cells<3, true><<<grid_for(s->n_cells, B), B, 0, st>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, G, cursor.p, L.bin_cells.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(st));
    L.grid = G;
    L.built = true;
    if (getenv("FDB_VERBOSE"))
        fprintf(stderr, "[fdb] locator: %d^%d bins, %d (cell, bin) pairs for %d cells\n", g, N, total, s->n_cells);
    return FDB_OK;
}

// device-side worker: locs_d column-major n_locs x N on the device, ids_d out
static int locate_device(fdb_space* s, int64_t n_locs, const double* locs_d, int32_t* ids_d) {
    FDB_TRY(build_locator(s));
    const GridDesc G = s->locator.grid;
    const int B = 128;
    if (s->M != s->N) return surface_locate(s, G, n_locs, locs_d, ids_d);
    if (s->M == 2)
        k_locate<2><<<grid_for(n_locs, B), B, 0, s->stream>>>(n_locs, locs_d, s->n_cells, s->n_nodes, s->verts_p, s->coords.p, G,
                                                              s->locator.bin_ptr.p, s->locator.bin_cells.p, ids_d);
    else
        k_locate<3><<<grid_for(n_locs, B), B, 0, s->stream>>>(n_locs, locs_d, s->n_cells, s->n_nodes, s->verts_p, s->coords.p, G,
                                                              s->locator.bin_ptr.p, s->locator.bin_cells.p, ids_d);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

int locate_host(fdb_space* s, int64_t n_locs, const double* locs, int32_t* ids) {
    FDB_CHECK(s && locs && ids && n_locs > 0, FDB_ERR_ARG, "fdb_locate: bad argument");
    FDB_CHECK(n_locs < (int64_t(1) << 31), FDB_ERR_UNSUPPORTED, "more than 2^31 locations");
    DevBuf<double> L;
    DevBuf<int32_t> I;
    FDB_TRY(L.alloc((size_t)n_locs * s->N));
    FDB_TRY(I.alloc((size_t)n_locs));
    FDB_CUDA(cudaMemcpyAsync(L.p, locs, sizeof(double) * n_locs * s->N, cudaMemcpyHostToDevice, s->stream));
    FDB_TRY(locate_device(s, n_locs, L.p, I.p));
    FDB_CUDA(cudaMemcpyAsync(ids, I.p, sizeof(int32_t) * n_locs, cudaMemcpyDeviceToHost, s->stream));
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    return FDB_OK;
}

int eval_pointwise_host(fdb_space* s, int64_t n_locs, const double* locs, int32_t* ids, int32_t* cols, double* vals) {
    FDB_CHECK(s && locs && cols && vals && n_locs > 0, FDB_ERR_ARG, "fdb_eval_pointwise: bad argument");
    FDB_CHECK(n_locs * s->nb < (int64_t(1) << 31), FDB_ERR_UNSUPPORTED, "n_locs * n_basis exceeds int32");
    DevBuf<double> L, V;
    DevBuf<int32_t> I, Cc;
    FDB_TRY(L.alloc((size_t)n_locs * s->N));
    FDB_TRY(I.alloc((size_t)n_locs));
    FDB_TRY(Cc.alloc((size_t)n_locs * s->nb));
    FDB_TRY(V.alloc((size_t)n_locs * s->nb));
    cudaStream_t st = s->stream;
    FDB_CUDA(cudaMemcpyAsync(L.p, locs, sizeof(double) * n_locs * s->N, cudaMemcpyHostToDevice, st));
    FDB_TRY(locate_device(s, n_locs, L.p, I.p));
    const int B = 128;
    if (s->M != s->N) FDB_TRY(surface_eval_pointwise(s, n_locs, L.p, I.p, Cc.p, V.p));
    else if (s->M == 2)
        k_eval_pointwise<2><<<grid_for(n_locs, B), B, 0, st>>>(n_locs, L.p, I.p, s->n_cells, s->n_nodes, s->verts_p, s->coords.p,
                                                               s->dofs.p, s->poly.p, Cc.p, V.p);
    else
        k_eval_pointwise<3><<<grid_for(n_locs, B), B, 0, st>>>(n_locs, L.p, I.p, s->n_cells, s->n_nodes, s->verts_p, s->coords.p,
                                                               s->dofs.p, s->poly.p, Cc.p, V.p);
    FDB_CUDA(cudaGetLastError());
    if (ids) FDB_CUDA(cudaMemcpyAsync(ids, I.p, sizeof(int32_t) * n_locs, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(cols, Cc.p, sizeof(int32_t) * n_locs * s->nb, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(vals, V.p, sizeof(double) * n_locs * s->nb, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    return FDB_OK;
}

int eval_areal_host(fdb_space* s, int n_sub, const double* incidence, int64_t capacity, int64_t* n_triplets, int32_t* rows,
                    int32_t* cols, double* vals, double* D) {
    FDB_CHECK(s && incidence && n_triplets && n_sub > 0, FDB_ERR_ARG, "fdb_eval_areal: bad argument");
    FDB_CHECK(s->M == s->N || (s->M == 2 && s->N == 3), FDB_ERR_UNSUPPORTED, "areal evaluation: unsupported mesh kind");
    const int64_t pairs = (int64_t)n_sub * s->n_cells;
    FDB_CHECK(pairs < (int64_t(1) << 31), FDB_ERR_UNSUPPORTED, "n_subdomains * n_cells exceeds int32");
    cudaStream_t st = s->stream;
    DevBuf<double> inc, integ, meas, Dd;
    DevBuf<int32_t> flag, pos;
    FDB_TRY(inc.alloc((size_t)pairs));
    FDB_TRY(flag.alloc((size_t)pairs + 1));
    FDB_TRY(pos.alloc((size_t)pairs + 1));
    FDB_TRY(integ.alloc((size_t)s->n_cells * s->nb));
    FDB_TRY(meas.alloc((size_t)s->n_cells));
    FDB_TRY(Dd.alloc((size_t)n_sub));
    FDB_CUDA(cudaMemcpyAsync(inc.p, incidence, sizeof(double) * pairs, cudaMemcpyHostToDevice, st));
    const int B = 128;
    k_incidence_flags<<<grid_for(pairs, 256), 256, 0, st>>>(n_sub, s->n_cells, inc.p, flag.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaMemsetAsync(flag.p + pairs, 0, sizeof(int32_t), st));
    FDB_CUDA(cudaMemcpyAsync(pos.p, flag.p, sizeof(int32_t) * (pairs + 1), cudaMemcpyDeviceToDevice, st));
    FDB_TRY(exclusive_scan(pos.p, pairs + 1, st));
    int32_t n_in = 0;
    FDB_CUDA(cudaMemcpyAsync(&n_in, pos.p + pairs, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    const int64_t nt = (int64_t)n_in * s->nb;
    *n_triplets = nt;
    if (!rows && !cols && !vals && !D) return FDB_OK;   // size query
    FDB_CHECK(rows && cols && vals && D, FDB_ERR_ARG, "fdb_eval_areal: null output array");
    FDB_CHECK(capacity >= nt, FDB_ERR_ARG, "fdb_eval_areal: output arrays too small");
    if (s->M != s->N) FDB_TRY(surface_cell_basis_integrals(s, integ.p, meas.p));
    else if (s->M == 2)
        k_cell_basis_integrals<2><<<grid_for(s->n_cells, B), B, 0, st>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p,
                                                                         s->poly.p, integ.p, meas.p);
    else
        k_cell_basis_integrals<3><<<grid_for(s->n_cells, B), B, 0, st>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p,
                                                                         s->poly.p, integ.p, meas.p);
    FDB_CUDA(cudaGetLastError());
    k_subdomain_measure<<<grid_for(n_sub, 64), 64, 0, st>>>(n_sub, s->n_cells, inc.p, meas.p, Dd.p);
    FDB_CUDA(cudaGetLastError());
    DevBuf<int32_t> r, c;
    DevBuf<double> v;
    FDB_TRY(r.alloc((size_t)std::max<int64_t>(nt, 1)));
    FDB_TRY(c.alloc((size_t)std::max<int64_t>(nt, 1)));
    FDB_TRY(v.alloc((size_t)std::max<int64_t>(nt, 1)));
    k_areal_fill<<<grid_for(pairs, 256), 256, 0, st>>>(n_sub, s->n_cells, s->nb, flag.p, pos.p, s->dofs.p, integ.p, Dd.p, r.p,
                                                        c.p, v.p);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaMemcpyAsync(rows, r.p, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(cols, c.p, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(vals, v.p, sizeof(double) * nt, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(D, Dd.p, sizeof(double) * n_sub, cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    return FDB_OK;
}

}  // namespace fdb
