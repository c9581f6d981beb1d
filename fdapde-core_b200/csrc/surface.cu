// Next-row N3: 2.5D manifold cells, Triangulation<2,3> (triangles embedded in R^3).
//
//   Simplex::initialize, manifold branch          geometry/simplex.h:189-193
//     J (3 x 2), J^+ = (J^T J)^-1 J^T (generalised inverse), measure = |J_0 x J_1| / 2
//   weak forms with g_i = (J^+)^T grad psi_i in R^3  operators/{laplacian,diffusion,advection,reaction}.h
// The pattern, scatter map and in-order segmented reduction are dimension independent (pattern.cu, assemble.cu); this
// file only supplies the per-cell kernels of the contribution-list path for N != M: local matrices, load vector,
// quadrature nodes and dof coordinates.  Surfaces are small next to the volumetric meshes of the hot path, so they do
// not get a fused plan.
#include <climits>

#include "local_matrix.cuh"

namespace fdb {

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

struct GeoS {
    double J[3][2];     // J[r][m]
    double invJ[2][3];  // invJ[m][r]
    double x0[3];
    double measure;
};

__device__ __forceinline__ void load_geometry_surface(int e, int n_cells, int n_nodes, const int32_t* __restrict__ verts,
                                                      const double* __restrict__ coords, GeoS& g) {
    double x[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int v = __ldg(verts + (size_t)k * n_cells + e);
#pragma unroll
        for (int r = 0; r < 3; ++r) x[k][r] = __ldg(coords + (size_t)r * n_nodes + v);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        g.x0[r] = x[0][r];
        g.J[r][0] = x[1][r] - x[0][r];
        g.J[r][1] = x[2][r] - x[0][r];
    }
    // G = J^T J = [a b; b d], G^-1 = [d -b; -b a] / det, J^+ = G^-1 J^T
    double a = 0, b = 0, d = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        a += g.J[r][0] * g.J[r][0];
        b += g.J[r][0] * g.J[r][1];
        d += g.J[r][1] * g.J[r][1];
    }
    const double det = a * d - b * b, invdet = 1.0 / det;
    const double gi[2][2] = {{d * invdet, -b * invdet}, {-b * invdet, a * invdet}};
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int r = 0; r < 3; ++r) g.invJ[m][r] = gi[m][0] * g.J[r][0] + gi[m][1] * g.J[r][1];
    const double cx = g.J[1][0] * g.J[2][1] - g.J[2][0] * g.J[1][1];
    const double cy = g.J[2][0] * g.J[0][1] - g.J[0][0] * g.J[2][1];
    const double cz = g.J[0][0] * g.J[1][1] - g.J[1][0] * g.J[0][1];
    g.measure = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
}

// one thread per cell; same quadrature loop and term order as local_matrix (local_matrix.cuh) with 3-vectors
template <int R, bool SYM>
__global__ void __launch_bounds__(128)
k_local_assemble_surface(int n_cells, int n_nodes, const int32_t* __restrict__ verts, const double* __restrict__ coords,
                         const FeTables* __restrict__ tab, OpCanon op, const int32_t* __restrict__ pos,
                         double* __restrict__ contrib) {
    constexpr int M = 2, N = 3, NB = nbasis(M, R), NQ = nquad(M, R), NE = nentries(M, R, SYM);
    __shared__ FeTables T;
    stage_tables(tab, &T);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    GeoS geo;
    load_geometry_surface(e, n_cells, n_nodes, verts, coords, geo);
    double acc[NE];
#pragma unroll
    for (int s = 0; s < NE; ++s) acc[s] = 0.0;
    const bool need_grad = op.has_lap | op.has_diff | op.has_adv;
#pragma unroll 1
    for (int q = 0; q < NQ; ++q) {
        const double wq = T.w[q];
        double g[NB][N], kg[NB][N], bg[NB], phi[NB];
        if (need_grad) {
#pragma unroll
            for (int i = 0; i < NB; ++i)
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double s = 0;
#pragma unroll
                    for (int m = 0; m < M; ++m) s += geo.invJ[m][r] * T.gref[(q * NB + i) * M + m];
                    g[i][r] = s;
                }
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) phi[i] = T.phi[q * NB + i];
        double cq = op.c;
        if (op.has_diff) {
            double K[N * N];
#pragma unroll
            for (int k = 0; k < N * N; ++k) K[k] = op.sv_diff ? op.Kp[((size_t)NQ * e + q) * (N * N) + k] : op.K[k];
#pragma unroll
            for (int j = 0; j < NB; ++j)
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double s = 0;
#pragma unroll
                    for (int c = 0; c < N; ++c) s += K[c * N + r] * g[j][c];
                    kg[j][r] = s;
                }
        }
        if (op.has_adv) {
            double bb[N];
#pragma unroll
            for (int r = 0; r < N; ++r) bb[r] = op.sv_adv ? op.bp[((size_t)NQ * e + q) * N + r] : op.b[r];
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                double s = 0;
#pragma unroll
                for (int r = 0; r < N; ++r) s += g[j][r] * bb[r];
                bg[j] = s;
            }
        }
        if (op.has_reac && op.sv_reac) cq = op.cp[(size_t)NQ * e + q];
        int s_idx = 0;
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int j = (SYM ? i : 0); j < NB; ++j) {
                double val = 0.0;
                if (op.has_lap) {
                    double d = 0;
#pragma unroll
                    for (int r = 0; r < N; ++r) d += g[i][r] * g[j][r];
                    val += op.s_lap * (-d);
                }
                if (op.has_diff) {
                    double d = 0;
#pragma unroll
                    for (int r = 0; r < N; ++r) d += g[i][r] * kg[j][r];
                    val += op.s_diff * (-d);
                }
                if (op.has_adv) val += op.s_adv * (phi[i] * bg[j]);
                if (op.has_reac) val += op.s_reac * (cq * phi[i] * phi[j]);
                acc[s_idx] += val * wq;
                ++s_idx;
            }
    }
#pragma unroll
    for (int s = 0; s < NE; ++s) contrib[pos[(size_t)s * n_cells + e]] = acc[s] * geo.measure;
}

template <int R>
__global__ void __launch_bounds__(128)
k_local_forcing_surface(int n_cells, int n_nodes, const int32_t* __restrict__ verts, const double* __restrict__ coords,
                        const FeTables* __restrict__ tab, const double* __restrict__ f_quad,
                        const int32_t* __restrict__ pos, double* __restrict__ contrib) {
    constexpr int NB = nbasis(2, R), NQ = nquad(2, R);
    __shared__ FeTables T;
    stage_tables(tab, &T);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    GeoS geo;
    load_geometry_surface(e, n_cells, n_nodes, verts, coords, geo);
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        double value = 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) value += (f_quad[(size_t)NQ * e + q] * T.phi[q * NB + i]) * T.w[q];
        contrib[pos[(size_t)i * n_cells + e]] = value * geo.measure;
    }
}

template <int R>
__global__ void k_quadrature_nodes_surface(int n_cells, int n_nodes, const int32_t* __restrict__ verts,
                                           const double* __restrict__ coords, const FeTables* __restrict__ tab,
                                           double* __restrict__ out) {
    constexpr int NQ = nquad(2, R);
    __shared__ FeTables T;
    stage_tables(tab, &T);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    GeoS geo;
    load_geometry_surface(e, n_cells, n_nodes, verts, coords, geo);
    const size_t rows = (size_t)n_cells * NQ;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double s = 0;
#pragma unroll
            for (int m = 0; m < 2; ++m) s += geo.J[r][m] * T.qn[q * 2 + m];
            out[(size_t)r * rows + (size_t)NQ * e + q] = s + geo.x0[r];
        }
}

__global__ void k_dof_coords_surface(int n_cells, int n_nodes, int n_dofs, int nb, int first_slot,
                                     const int32_t* __restrict__ verts, const int32_t* __restrict__ dofs,
                                     const double* __restrict__ coords, const FeTables* __restrict__ tab,
                                     const int32_t* __restrict__ first, double* __restrict__ out) {
    __shared__ FeTables T;
    stage_tables(tab, &T);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    GeoS geo;
    load_geometry_surface(e, n_cells, n_nodes, verts, coords, geo);
    for (int j = first_slot; j < nb; ++j) {
        const int d = dofs[(size_t)j * n_cells + e];
        if (first[d] != e) continue;
        if (j <= 2) {
            const int v = verts[(size_t)j * n_cells + e];
            for (int r = 0; r < 3; ++r) out[(size_t)r * n_dofs + d] = coords[(size_t)r * n_nodes + v];
            continue;
        }
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int m = 0; m < 2; ++m) s += geo.J[r][m] * T.refn[j * 2 + m];
            out[(size_t)r * n_dofs + d] = s + geo.x0[r];
        }
    }
}

// ---- point location and basis evaluation on surfaces (evaluate.cu dispatches here when N != M) ---------------------------
// Simplex::contains for manifold cells (simplex.h:115-128): the point must lie on the supporting plane -- distance to
// its projection B B^T (x - p) + p with the orthonormal basis of HyperPlane<2,3> (hyperplane.h:56-62, 92-99) at most
// 10 eps -- and have barycentric coordinates z >= -10 eps, z(1..2) = J^+ (x - v0).
__device__ __forceinline__ bool surface_contains(const GeoS& g, const double* p) {
    const double meps = 10 * 2.220446049250313e-16;
    double b0[3], b1[3], d[3], n0 = 0, n1 = 0, wb = 0, bb = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) { b0[r] = g.J[r][0]; n0 += b0[r] * b0[r]; d[r] = p[r] - g.x0[r]; }
    n0 = sqrt(n0);
#pragma unroll
    for (int r = 0; r < 3; ++r) b0[r] /= n0;
#pragma unroll
    for (int r = 0; r < 3; ++r) { wb += g.J[r][1] * b0[r]; bb += b0[r] * b0[r]; }
#pragma unroll
    for (int r = 0; r < 3; ++r) { b1[r] = g.J[r][1] - wb / bb * b0[r]; n1 += b1[r] * b1[r]; }
    n1 = sqrt(n1);
#pragma unroll
    for (int r = 0; r < 3; ++r) b1[r] /= n1;
    double c0 = 0, c1 = 0, dist = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) { c0 += b0[r] * d[r]; c1 += b1[r] * d[r]; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double pr = (b0[r] * c0 + b1[r] * c1) + g.x0[r];
        dist += (p[r] - pr) * (p[r] - pr);
    }
    if (sqrt(dist) > meps) return false;
    double sum = 0;
    bool in = true;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        double t = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) t += g.invJ[m][r] * d[r];
        sum += t;
        in = in && !(t < -meps);
    }
    return in && !((1 - sum) < -meps);
}

template <bool FILL>
__global__ void k_bin_cells_surface(int n_cells, int n_nodes, const int32_t* __restrict__ verts,
                                    const double* __restrict__ coords, GridDesc G, int32_t* __restrict__ counter,
                                    int32_t* __restrict__ bin_cells) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    int b0[3], b1[3];
    for (int d = 0; d < 3; ++d) {
        double lo = 0, hi = 0;
        for (int k = 0; k < 3; ++k) {
            const double x = coords[(size_t)d * n_nodes + verts[(size_t)k * n_cells + e]];
            lo = k == 0 ? x : fmin(lo, x);
            hi = k == 0 ? x : fmax(hi, x);
        }
        b0[d] = bin_of(G, d, lo - G.eps[d]);
        b1[d] = bin_of(G, d, hi + G.eps[d]);
    }
    for (int c = b0[2]; c <= b1[2]; ++c)
        for (int b = b0[1]; b <= b1[1]; ++b)
            for (int a = b0[0]; a <= b1[0]; ++a) {
                const int bin = (c * G.g[1] + b) * G.g[0] + a;
                const int slot = atomicAdd(&counter[bin], 1);
                if (FILL) bin_cells[slot] = e;
            }
}

__global__ void k_locate_surface(int64_t n_locs, const double* __restrict__ locs, int n_cells, int n_nodes,
                                 const int32_t* __restrict__ verts, const double* __restrict__ coords, GridDesc G,
                                 const int32_t* __restrict__ bin_ptr, const int32_t* __restrict__ bin_cells,
                                 int32_t* __restrict__ ids) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_locs) return;
    double p[3];
    int bin = 0;
    for (int d = 2; d >= 0; --d) {
        p[d] = locs[(size_t)d * n_locs + i];
        bin = bin * G.g[d] + bin_of(G, d, p[d]);
    }
    int best = INT_MAX;
    for (int t = bin_ptr[bin]; t < bin_ptr[bin + 1]; ++t) {
        const int e = bin_cells[t];
        if (e >= best) continue;
        GeoS g;
        load_geometry_surface(e, n_cells, n_nodes, verts, coords, g);
        if (surface_contains(g, p)) best = e;
    }
    ids[i] = best == INT_MAX ? -1 : best;
}

__global__ void k_eval_pointwise_surface(int64_t n_locs, const double* __restrict__ locs, const int32_t* __restrict__ ids,
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
    GeoS g;
    load_geometry_surface(e, n_cells, n_nodes, verts, coords, g);
    double xi[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        double t = 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) t += g.invJ[m][r] * (locs[(size_t)r * n_locs + i] - g.x0[r]);
        xi[m] = t;
    }
    for (int h = 0; h < nb; ++h) {
        cols[i * nb + h] = dofs[(size_t)h * n_cells + e];
        vals[i * nb + h] = poly_eval(P, h, xi);
    }
}

__global__ void k_cell_basis_integrals_surface(int n_cells, int n_nodes, const int32_t* __restrict__ verts,
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
    GeoS g;
    load_geometry_surface(e, n_cells, n_nodes, verts, coords, g);
    const int nb = P.nb;
    for (int h = 0; h < nb; ++h) {
        double value = 0;
        for (int q = 0; q < T.nq; ++q) {
            double p[3], xi[2];
#pragma unroll
            for (int r = 0; r < 3; ++r) p[r] = (g.J[r][0] * T.qn[q * 2] + g.J[r][1] * T.qn[q * 2 + 1]) + g.x0[r];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                double t = 0;
#pragma unroll
                for (int r = 0; r < 3; ++r) t += g.invJ[m][r] * (p[r] - g.x0[r]);
                xi[m] = t;
            }
            value += poly_eval(P, h, xi) * T.w[q];
        }
        integ[(size_t)e * nb + h] = value * g.measure;
    }
    meas[e] = g.measure;
}

int surface_bin_cells(fdb_space* s, const GridDesc& G, int32_t* counter, int32_t* bin_cells, bool fill) {
    const int B = 128;
    if (fill) k_bin_cells_surface<true><<<grid_for(s->n_cells, B), B, 0, s->stream>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, G, counter, bin_cells);
    else k_bin_cells_surface<false><<<grid_for(s->n_cells, B), B, 0, s->stream>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, G, counter, bin_cells);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}
int surface_locate(fdb_space* s, const GridDesc& G, int64_t n_locs, const double* locs_d, int32_t* ids_d) {
    k_locate_surface<<<grid_for(n_locs, 128), 128, 0, s->stream>>>(n_locs, locs_d, s->n_cells, s->n_nodes, s->verts_p, s->coords.p, G,
                                                                   s->locator.bin_ptr.p, s->locator.bin_cells.p, ids_d);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}
int surface_eval_pointwise(fdb_space* s, int64_t n_locs, const double* locs_d, const int32_t* ids_d, int32_t* cols, double* vals) {
    k_eval_pointwise_surface<<<grid_for(n_locs, 128), 128, 0, s->stream>>>(n_locs, locs_d, ids_d, s->n_cells, s->n_nodes, s->verts_p,
                                                                           s->coords.p, s->dofs.p, s->poly.p, cols, vals);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}
int surface_cell_basis_integrals(fdb_space* s, double* integ, double* meas) {
    k_cell_basis_integrals_surface<<<grid_for(s->n_cells, 128), 128, 0, s->stream>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p,
                                                                                     s->tab.p, s->poly.p, integ, meas);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

// ---- launchers (called from assemble.cu when N != M) -------------------------------------------------------------------
int surface_local_assemble(fdb_space* s, const Pattern& P, const OpCanon& op, double* contrib) {
    const int B = 128;
    const unsigned G = grid_for(s->n_cells, B);
#define FDB_LAUNCH_S(RR, SS)                                                                                           \
    k_local_assemble_surface<RR, SS><<<G, B, 0, s->stream>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, op, \
                                                             P.pos.p, contrib)
    if (s->R == 1 && P.symmetric) FDB_LAUNCH_S(1, true);
    else if (s->R == 1) FDB_LAUNCH_S(1, false);
    else if (P.symmetric) FDB_LAUNCH_S(2, true);
    else FDB_LAUNCH_S(2, false);
#undef FDB_LAUNCH_S
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

int surface_local_forcing(fdb_space* s, const double* f_quad, const int32_t* pos, double* contrib) {
    const int B = 128;
    const unsigned G = grid_for(s->n_cells, B);
    if (s->R == 1)
        k_local_forcing_surface<1><<<G, B, 0, s->stream>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, f_quad, pos, contrib);
    else
        k_local_forcing_surface<2><<<G, B, 0, s->stream>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, f_quad, pos, contrib);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

int surface_quadrature_nodes(fdb_space* s, double* out) {
    const int B = 128;
    const unsigned G = grid_for(s->n_cells, B);
    if (s->R == 1) k_quadrature_nodes_surface<1><<<G, B, 0, s->stream>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, out);
    else k_quadrature_nodes_surface<2><<<G, B, 0, s->stream>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, out);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

int surface_dof_coords(fdb_space* s, int first_slot, const int32_t* first, double* out) {
    k_dof_coords_surface<<<grid_for(s->n_cells, 128), 128, 0, s->stream>>>(s->n_cells, s->n_nodes, s->n_dofs, s->nb, first_slot,
                                                                           s->verts_p, s->dofs.p, s->coords.p, s->tab.p, first, out);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

}  // namespace fdb
