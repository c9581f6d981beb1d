// Next-row N3: 2.5D manifold cells, Triangulation<2,3> (triangles embedded in R^3).
//
//   Simplex::initialize, manifold branch          geometry/simplex.h:189-193
//     J (3 x 2), J^+ = (J^T J)^-1 J^T (generalised inverse), measure = |J_0 x J_1| / 2
//   weak forms with g_i = (J^+)^T grad psi_i in R^3  operators/{laplacian,diffusion,advection,reaction}.h
// The pattern, scatter map and in-order segmented reduction are dimension independent (pattern.cu, assemble.cu); this
// file only supplies the per-cell kernels of the contribution-list path for N != M: local matrices, load vector,
// quadrature nodes and dof coordinates.  Surfaces are small next to the volumetric meshes of the hot path, so they do
// not get a fused plan.
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
