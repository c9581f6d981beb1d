// Device-side building blocks shared by the assembly kernels: cell geometry and the local element matrix.
//
//   Simplex::initialize (J, J^-1, measure)      geometry/simplex.h:184-195
//   Integrator::integrate_weak_form             utils/integration/integrator.h:93-106
//   weak forms                                  operators/{laplacian,diffusion,advection,reaction,dt}.h
#pragma once
#include "common.cuh"

namespace fdb {

// canonical form of the operator expression tree: at most one term of each kind
struct OpCanon {
    int has_lap, has_diff, has_adv, has_reac;
    int sv_diff, sv_adv, sv_reac;
    double s_lap, s_diff, s_adv, s_reac;
    double K[MAX_D * MAX_D];  // column-major N x N
    double b[MAX_D];
    double c;
    double wsum;       // sum of the quadrature weights, accumulated left to right like the quadrature loop
    double lap_k0;     // s_lap * wsum / M!  (lean P1 stiffness path)
    const double* Kp;  // space-varying coefficient rows (device), row nq*e+q
    const double* bp;
    const double* cp;
};

constexpr __host__ __device__ int nbasis(int M, int R) { return R == 1 ? M + 1 : (M + 1) * (M + 2) / 2; }
constexpr __host__ __device__ int nquad(int M, int R) { return M == 2 ? (R == 1 ? 3 : 6) : (R == 1 ? 4 : 5); }
constexpr __host__ __device__ int nentries(int M, int R, bool sym) {
    return sym ? nbasis(M, R) * (nbasis(M, R) + 1) / 2 : nbasis(M, R) * nbasis(M, R);
}

template <int M> struct Geo {
    double invJ[M][M];  // invJ[m][r]
    double J[M][M];     // J[r][m]
    double x0[M];
    double measure;
};

// J, J^-1 (adjugate / determinant, the cofactor expansion of a fixed-size inverse) and measure from the vertices
template <int M>
__device__ __forceinline__ void finish_geometry(const double (&x)[M + 1][M], Geo<M>& g) {
#pragma unroll
    for (int r = 0; r < M; ++r) {
        g.x0[r] = x[0][r];
#pragma unroll
        for (int m = 0; m < M; ++m) g.J[r][m] = x[m + 1][r] - x[0][r];
    }
    if constexpr (M == 2) {
        double det = g.J[0][0] * g.J[1][1] - g.J[1][0] * g.J[0][1];
        double invdet = 1.0 / det;
        g.invJ[0][0] = g.J[1][1] * invdet;
        g.invJ[1][0] = -g.J[1][0] * invdet;
        g.invJ[0][1] = -g.J[0][1] * invdet;
        g.invJ[1][1] = g.J[0][0] * invdet;
        g.measure = fabs(det) / 2;
    } else {
        // adjugate / determinant, same cofactor expansion as a fixed-size 3x3 inverse
        double c[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                c[i][j] = g.J[i1][j1] * g.J[i2][j2] - g.J[i1][j2] * g.J[i2][j1];
            }
        double det = c[0][0] * g.J[0][0] + c[1][0] * g.J[1][0] + c[2][0] * g.J[2][0];
        double invdet = 1.0 / det;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) g.invJ[i][j] = c[j][i] * invdet;
        g.measure = fabs(det) / 6;
    }
}

// vertex coordinates of cell e from the struct-of-arrays copies (coalesced vertex ids, gathered coordinates)
template <int M>
__device__ __forceinline__ void gather_vertices(int e, int n_cells, int n_nodes, const int32_t* __restrict__ verts,
                                                const double* __restrict__ coords, double (&x)[M + 1][M]) {
    int v[M + 1];
#pragma unroll
    for (int k = 0; k <= M; ++k) v[k] = __ldg(verts + (size_t)k * n_cells + e);
#pragma unroll
    for (int k = 0; k <= M; ++k)
#pragma unroll
        for (int r = 0; r < M; ++r) x[k][r] = __ldg(coords + (size_t)r * n_nodes + v[k]);
}

template <int M>
__device__ __forceinline__ void load_geometry(int e, int n_cells, int n_nodes, const int32_t* __restrict__ verts,
                                              const double* __restrict__ coords, Geo<M>& g) {
    double x[M + 1][M];
    gather_vertices<M>(e, n_cells, n_nodes, verts, coords, x);
    finish_geometry<M>(x, g);
}

// same geometry from block-major vertex ids (M+1 consecutive ints) and the packed coordinate copy
// (3D: 4 doubles per node = one 32-byte sector, 2D: 2 doubles per node)
template <int M>
struct VertexIds { int v[M + 1]; };

template <int M>
__device__ __forceinline__ VertexIds<M> load_vertex_ids(const int32_t* __restrict__ vp) {
    VertexIds<M> r;
    if constexpr (M == 3) {
        const int4 q = __ldg(reinterpret_cast<const int4*>(vp));
        r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w;
    } else {
#pragma unroll
        for (int k = 0; k <= M; ++k) r.v[k] = __ldg(vp + k);
    }
    return r;
}

template <int M>
__device__ __forceinline__ void gather_coords_packed(const VertexIds<M>& id, const double* __restrict__ pk,
                                                     double (&x)[M + 1][M]) {
#pragma unroll
    for (int k = 0; k <= M; ++k) {
        if constexpr (M == 3) {
            const double2 a = __ldg(reinterpret_cast<const double2*>(pk + (size_t)id.v[k] * 4));
            const double2 c = __ldg(reinterpret_cast<const double2*>(pk + (size_t)id.v[k] * 4) + 1);
            x[k][0] = a.x; x[k][1] = a.y; x[k][2] = c.x;
        } else {
            const double2 a = __ldg(reinterpret_cast<const double2*>(pk + (size_t)id.v[k] * 2));
            x[k][0] = a.x; x[k][1] = a.y;
        }
    }
}

template <int M>
__device__ __forceinline__ void gather_vertices_packed(const int32_t* __restrict__ vp, const double* __restrict__ pk,
                                                       double (&x)[M + 1][M]) {
    gather_coords_packed<M>(load_vertex_ids<M>(vp), pk, x);
}

// psi_h(xi): sum of coefficient * monomial, monomials ascending (multivariate_polynomial.h:111-145,209-213); powers
// are at most 2, so repeated multiplication reproduces std::pow exactly
__device__ __forceinline__ double poly_eval(const PolyTables& P, int h, const double* xi) {
    double v = 0;
    for (int m = 0; m < P.nb; ++m) {
        const int* e = P.ex + m * P.M;
        double mono = e[0] == 0 ? 1.0 : (e[0] == 1 ? xi[0] : xi[0] * xi[0]);
        for (int k = 1; k < P.M; ++k)
            if (e[k] != 0) mono = (e[k] == 1 ? xi[k] : xi[k] * xi[k]) * mono;
        const double t = P.coef[h * P.nb + m] * mono;
        v = (m == 0) ? t : t + v;
    }
    return v;
}

__device__ __forceinline__ int bin_of(const GridDesc& G, int d, double x) {
    int b = (int)floor((x - G.lo[d]) * G.inv_h[d]);
    return b < 0 ? 0 : (b >= G.g[d] ? G.g[d] - 1 : b);
}

__device__ __forceinline__ void stage_tables(const FeTables* __restrict__ tab, FeTables* sm) {
    const int words = sizeof(FeTables) / sizeof(int);
    const int* src = reinterpret_cast<const int*>(tab);
    int* dst = reinterpret_cast<int*>(sm);
    for (int k = threadIdx.x; k < words; k += blockDim.x) dst[k] = src[k];
    __syncthreads();
}

// The NE local entries of cell e, in emission-slot order (symmetric: pairs i <= j; else i outer, j inner).
// LAP = true is the lean instantiation for operators made of the Laplacian only (the stiffness matrix): with P1
// elements the integrand is constant over the cell, so the quadrature loop collapses to one multiplication by the
// (left-to-right) sum of the weights.
template <int M, int R, bool SYM, bool LAP>
__device__ __forceinline__ void local_matrix(const Geo<M>& geo, const FeTables& T, const OpCanon& op, int e,
                                             double (&acc)[nentries(M, R, SYM)]) {
    constexpr int NB = nbasis(M, R), NQ = nquad(M, R), NE = nentries(M, R, SYM);
    double g[NB][M];
    if constexpr (R == 1) {  // constant gradients: evaluate once
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int r = 0; r < M; ++r) {
                double s = 0;
#pragma unroll
                for (int m = 0; m < M; ++m) s += geo.invJ[m][r] * T.gref[i * M + m];
                g[i][r] = s;
            }
    }
    if constexpr (LAP && R == 1) {
        const double f = op.wsum * geo.measure;
        int s_idx = 0;
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int j = (SYM ? i : 0); j < NB; ++j) {
                double d = 0;
#pragma unroll
                for (int r = 0; r < M; ++r) d += g[i][r] * g[j][r];
                acc[s_idx++] = (op.s_lap * (-d)) * f;
            }
        return;
    }
#pragma unroll
    for (int s = 0; s < NE; ++s) acc[s] = 0.0;
    const bool need_grad = LAP || (op.has_lap | op.has_diff | op.has_adv);
#pragma unroll 1
    for (int q = 0; q < NQ; ++q) {
        const double wq = T.w[q];
        if constexpr (R != 1) {
            if (need_grad) {
#pragma unroll
                for (int i = 0; i < NB; ++i)
#pragma unroll
                    for (int r = 0; r < M; ++r) {
                        double s = 0;
#pragma unroll
                        for (int m = 0; m < M; ++m) s += geo.invJ[m][r] * T.gref[(q * NB + i) * M + m];
                        g[i][r] = s;
                    }
            }
        }
        if constexpr (LAP) {
            int s_idx = 0;
#pragma unroll
            for (int i = 0; i < NB; ++i)
#pragma unroll
                for (int j = (SYM ? i : 0); j < NB; ++j) {
                    double d = 0;
#pragma unroll
                    for (int r = 0; r < M; ++r) d += g[i][r] * g[j][r];
                    acc[s_idx++] += (op.s_lap * (-d)) * wq;
                }
        } else {
            double phi[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) phi[i] = T.phi[q * NB + i];
            double kg[NB][M];  // K g_j
            double bg[NB];     // g_j . b
            double cq = op.c;
            if (op.has_diff) {
                double K[M * M];
                if (op.sv_diff) {
                    const double* kp = op.Kp + ((size_t)NQ * e + q) * (M * M);
#pragma unroll
                    for (int k = 0; k < M * M; ++k) K[k] = kp[k];
                } else {
#pragma unroll
                    for (int k = 0; k < M * M; ++k) K[k] = op.K[k];
                }
#pragma unroll
                for (int j = 0; j < NB; ++j)
#pragma unroll
                    for (int r = 0; r < M; ++r) {
                        double s = 0;
#pragma unroll
                        for (int c = 0; c < M; ++c) s += K[c * M + r] * g[j][c];
                        kg[j][r] = s;
                    }
            }
            if (op.has_adv) {
                double bb[M];
                if (op.sv_adv) {
                    const double* bp = op.bp + ((size_t)NQ * e + q) * M;
#pragma unroll
                    for (int r = 0; r < M; ++r) bb[r] = bp[r];
                } else {
#pragma unroll
                    for (int r = 0; r < M; ++r) bb[r] = op.b[r];
                }
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    double s = 0;
#pragma unroll
                    for (int r = 0; r < M; ++r) s += g[j][r] * bb[r];
                    bg[j] = s;
                }
            }
            if (op.has_reac && op.sv_reac) cq = op.cp[(size_t)NQ * e + q];
            int s_idx = 0;
#pragma unroll
            for (int i = 0; i < NB; ++i) {
#pragma unroll
                for (int j = (SYM ? i : 0); j < NB; ++j) {
                    double val = 0.0;
                    if (op.has_lap) {
                        double d = 0;
#pragma unroll
                        for (int r = 0; r < M; ++r) d += g[i][r] * g[j][r];
                        val += op.s_lap * (-d);
                    }
                    if (op.has_diff) {
                        double d = 0;
#pragma unroll
                        for (int r = 0; r < M; ++r) d += g[i][r] * kg[j][r];
                        val += op.s_diff * (-d);
                    }
                    if (op.has_adv) val += op.s_adv * (phi[i] * bg[j]);
                    if (op.has_reac) val += op.s_reac * (cq * phi[i] * phi[j]);
                    acc[s_idx] += val * wq;
                    ++s_idx;
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < NE; ++s) acc[s] *= geo.measure;
}

// Lean path for the stiffness matrix of P1 elements (operator = scale * Laplacian, R == 1).  The P1 reference
// gradients are the exact constants (-1,..,-1), e_1, .., e_M, so with A = adj(J) (unscaled cofactors)
//   g_k = A[:, k-1] / det (k >= 1),  g_0 = -(g_1 + .. + g_M),
//   a_ij = s * (-(g_i . g_j)) * (sum_q w_q) * |det| / M!  =  (A_i . A_j) * ( -s * wsum / (M! |det|) ).
// One division per cell, no J^-1, no table reads.  op.lap_k0 = s * wsum / M!.
template <int M, bool SYM>
__device__ __forceinline__ void p1_laplacian_matrix(const double (&x)[M + 1][M], double lap_k0,
                                                    double (&acc)[nentries(M, 1, SYM)]) {
    constexpr int NB = M + 1;
    double J[M][M];
#pragma unroll
    for (int r = 0; r < M; ++r)
#pragma unroll
        for (int m = 0; m < M; ++m) J[r][m] = x[m + 1][r] - x[0][r];
    double G[NB][M];  // unscaled physical gradients of the vertices 1..M (row 0 unused)
    double det;
    if constexpr (M == 2) {
        det = J[0][0] * J[1][1] - J[1][0] * J[0][1];
        G[1][0] = J[1][1]; G[1][1] = -J[0][1];   // g_1 = row 0 of adj(J) = invJ[0][:] * det
        G[2][0] = -J[1][0]; G[2][1] = J[0][0];   // g_2 = row 1 of adj(J)
    } else {
        double c[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                c[i][j] = J[i1][j1] * J[i2][j2] - J[i1][j2] * J[i2][j1];
            }
        det = c[0][0] * J[0][0] + c[1][0] * J[1][0] + c[2][0] * J[2][0];
        // invJ[m][r] = c[r][m] / det  and  g_k[r] = invJ[k-1][r]
#pragma unroll
        for (int k = 1; k <= 3; ++k)
#pragma unroll
            for (int r = 0; r < 3; ++r) G[k][r] = c[r][k - 1];
    }
    // Entries among the vertices 1..M are dot products of adjugate rows; those involving vertex 0 follow from the
    // zero row sums of the stiffness matrix (g_0 = -(g_1 + .. + g_M)):  d_0j = -(d_1j + .. + d_Mj),
    // d_00 = -(d_01 + .. + d_0M).  6 instead of 10 dot products on tetrahedra.
    double D[NB][NB];
#pragma unroll
    for (int i = 1; i < NB; ++i)
#pragma unroll
        for (int j = i; j < NB; ++j) {
            double d = G[i][0] * G[j][0];
#pragma unroll
            for (int r = 1; r < M; ++r) d += G[i][r] * G[j][r];
            D[i][j] = d;
            D[j][i] = d;
        }
#pragma unroll
    for (int j = 1; j < NB; ++j) {
        double s = D[1][j];
#pragma unroll
        for (int k = 2; k < NB; ++k) s += D[k][j];
        D[0][j] = -s;
        D[j][0] = -s;
    }
    {
        double s = D[0][1];
#pragma unroll
        for (int k = 2; k < NB; ++k) s += D[0][k];
        D[0][0] = -s;
    }
    const double f = -lap_k0 / fabs(det);
    int s_idx = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i)
#pragma unroll
        for (int j = (SYM ? i : 0); j < NB; ++j) acc[s_idx++] = D[i][j] * f;
}

// ---- reference-tensor form of constant-coefficient operators ---------------------------------------------------------------
// With g_i = J^-T grad psi_i every term of the weak form factors into a per-cell M x M (or M-vector, scalar) weight and a
// constant tensor of the reference element, contracted over the quadrature rule once on the host:
//   -(g_i . g_j)    -> -sum_mn (J^-1 J^-T)_mn     T^mn_ij,   T^mn_ij = sum_q w_q d_m psi_i(p_q) d_n psi_j(p_q)
//   -(g_i . K g_j)  -> -sum_mn (J^-1 K J^-T)_mn   T^mn_ij
//   psi_i (g_j . b) ->  sum_n  (J^-1 b)_n          A^n_ij,    A^n_ij  = sum_q w_q psi_i(p_q) d_n psi_j(p_q)
//   c psi_i psi_j   ->  c                          R_ij,      R_ij    = sum_q w_q psi_i(p_q) psi_j(p_q)
// (the same quadrature formula as integrate_weak_form, integrator.h:93-106, re-associated): M^2 + M + 1 FMAs per entry
// instead of a quadrature loop with per-point gradients.  Table row of (i, j): [T^mn (M^2) | A^n (M) | R | pad], read as
// warp broadcasts from shared memory.
constexpr __host__ __device__ int tens_stride(int M) { return M == 2 ? 8 : 14; }

// per-cell weights of the reference tensors (already multiplied by the measure)
template <int M> struct TensWeights { double W[M * M], beta[M], gamma; };

template <int M>
__device__ __forceinline__ void tens_weights(const Geo<M>& geo, const OpCanon& op, TensWeights<M>& w) {
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
        for (int n = 0; n < M; ++n) {
            double acc = 0;
            if (op.has_lap) {
                double d = 0;
#pragma unroll
                for (int r = 0; r < M; ++r) d += geo.invJ[m][r] * geo.invJ[n][r];
                acc += op.s_lap * (-d);
            }
            if (op.has_diff) {
                double d = 0;
#pragma unroll
                for (int r = 0; r < M; ++r) {
                    double kr = 0;  // (K J^-T)_rn = sum_c K(r, c) invJ[n][c]
#pragma unroll
                    for (int c = 0; c < M; ++c) kr += op.K[c * M + r] * geo.invJ[n][c];
                    d += geo.invJ[m][r] * kr;
                }
                acc += op.s_diff * (-d);
            }
            w.W[m * M + n] = acc * geo.measure;
        }
#pragma unroll
    for (int n = 0; n < M; ++n) {
        double d = 0;
        if (op.has_adv) {
#pragma unroll
            for (int r = 0; r < M; ++r) d += geo.invJ[n][r] * op.b[r];
        }
        w.beta[n] = op.has_adv ? op.s_adv * d * geo.measure : 0.0;
    }
    w.gamma = op.has_reac ? op.s_reac * op.c * geo.measure : 0.0;
}

// how a kernel evaluates the local matrix
constexpr int MODE_LEAN = 0;       // P1 elements, operator = scale * Laplacian: closed form, no tables
constexpr int MODE_TENSOR = 1;     // constant coefficients: reference-tensor form, every term
constexpr int MODE_QUAD = 2;       // space-varying coefficients: quadrature loop
constexpr int MODE_TENS_LAP = 3;   // operator = scale * Laplacian (P2 stiffness): W = J^-1 J^-T is symmetric, so the table row
                                   // holds the M(M+1)/2 symmetrised tensors T^mn + T^nm (m < n), T^mm
constexpr int MODE_TENS_REAC = 4;  // operator = constant reaction (mass matrix): one tensor R_ij, weight c |det| / M!
constexpr __host__ __device__ bool is_tensor_mode(int mode) { return mode == MODE_TENSOR || mode == MODE_TENS_LAP || mode == MODE_TENS_REAC; }
// doubles per table row (i, j) in each tensor mode, and the offset of each mode's table inside fdb_space::tens
constexpr __host__ __device__ int tens_stride_of(int M, int mode) {
    return mode == MODE_TENS_LAP ? (M == 2 ? 4 : 6) : (mode == MODE_TENS_REAC ? 1 : tens_stride(M));
}
constexpr __host__ __device__ int tens_offset_of(int M, int nb, int mode) {
    return mode == MODE_TENS_LAP ? nb * nb * tens_stride(M)
                                 : (mode == MODE_TENS_REAC ? nb * nb * (tens_stride(M) + tens_stride_of(M, MODE_TENS_LAP)) : 0);
}
constexpr __host__ __device__ int tens_total(int M, int nb) {
    return nb * nb * (tens_stride(M) + tens_stride_of(M, MODE_TENS_LAP) + 1);
}

// entry (i, j) of the local matrix against its table row (same address across the warp: broadcast loads)
template <int M, int MODE = MODE_TENSOR>
__device__ __forceinline__ double tens_entry(const double* __restrict__ tab, int ij, const TensWeights<M>& w) {
    if constexpr (MODE == MODE_TENS_REAC) {
        return w.gamma * tab[ij];
    } else if constexpr (MODE == MODE_TENS_LAP) {
        constexpr int TS = tens_stride_of(M, MODE_TENS_LAP);
        const double2* t2 = reinterpret_cast<const double2*>(tab + ij * TS);
        if constexpr (M == 2) {
            const double2 a = t2[0], b = t2[1];   // T00, T01 + T10, T11, pad
            double v = w.W[0] * a.x;
            v = fma(w.W[1], a.y, v);
            v = fma(w.W[3], b.x, v);
            return v;
        } else {
            const double2 a = t2[0], b = t2[1], c = t2[2];   // T00, T01+T10, T02+T20, T11, T12+T21, T22
            double v = w.W[0] * a.x;
            v = fma(w.W[1], a.y, v);
            v = fma(w.W[2], b.x, v);
            v = fma(w.W[4], b.y, v);
            v = fma(w.W[5], c.x, v);
            v = fma(w.W[8], c.y, v);
            return v;
        }
    } else {
        constexpr int TS = tens_stride(M);
        const double2* t2 = reinterpret_cast<const double2*>(tab + ij * TS);
        double t[TS];
#pragma unroll
        for (int k = 0; k < TS / 2; ++k) { const double2 q = t2[k]; t[2 * k] = q.x; t[2 * k + 1] = q.y; }
        double v = w.gamma * t[M * M + M];
#pragma unroll
        for (int n = 0; n < M; ++n) v += w.beta[n] * t[M * M + n];
#pragma unroll
        for (int k = 0; k < M * M; ++k) v += w.W[k] * t[k];
        return v;
    }
}

// per-cell weights a tensor mode actually reads (the split phase 1 of the fused kernel passes them through shared memory)
constexpr __host__ __device__ int tens_nw(int M, int mode) {
    return mode == MODE_TENS_REAC ? 1 : (mode == MODE_TENS_LAP ? M * (M + 1) / 2 : M * M + M + 1);
}
template <int M, int MODE>
__device__ __forceinline__ void tens_pack(const TensWeights<M>& w, double* out, int stride) {
    if constexpr (MODE == MODE_TENS_REAC) {
        out[0] = w.gamma;
    } else if constexpr (MODE == MODE_TENS_LAP) {
        int k = 0;
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
            for (int n = m; n < M; ++n) out[(k++) * stride] = w.W[m * M + n];
    } else {
#pragma unroll
        for (int k = 0; k < M * M; ++k) out[k * stride] = w.W[k];
#pragma unroll
        for (int n = 0; n < M; ++n) out[(M * M + n) * stride] = w.beta[n];
        out[(M * M + M) * stride] = w.gamma;
    }
}
template <int M, int MODE>
__device__ __forceinline__ void tens_unpack(const double* in, int stride, TensWeights<M>& w) {
    if constexpr (MODE == MODE_TENS_REAC) {
        w.gamma = in[0];
    } else if constexpr (MODE == MODE_TENS_LAP) {
        int k = 0;
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
            for (int n = m; n < M; ++n) w.W[m * M + n] = in[(k++) * stride];
    } else {
#pragma unroll
        for (int k = 0; k < M * M; ++k) w.W[k] = in[k * stride];
#pragma unroll
        for (int n = 0; n < M; ++n) w.beta[n] = in[(M * M + n) * stride];
        w.gamma = in[(M * M + M) * stride];
    }
}

// stages the nb^2 table rows of a space into shared memory
__device__ __forceinline__ void stage_tensor_table(const double* __restrict__ tens, double* sm, int count) {
    for (int k = threadIdx.x; k < count; k += blockDim.x) sm[k] = tens[k];
    __syncthreads();
}

// local matrix of one cell from its vertex coordinates
template <int M, int R, bool SYM, bool LAP>
__device__ __forceinline__ void cell_matrix(const double (&x)[M + 1][M], const FeTables& T, const OpCanon& op, int e,
                                            double (&acc)[nentries(M, R, SYM)]) {
    if constexpr (LAP && R == 1) {
        p1_laplacian_matrix<M, SYM>(x, op.lap_k0, acc);
    } else {
        Geo<M> geo;
        finish_geometry<M>(x, geo);
        local_matrix<M, R, SYM, LAP>(geo, T, op, e, acc);
    }
}

}  // namespace fdb
