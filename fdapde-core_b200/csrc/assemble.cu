// K3/K4/K5/K6: local assembly in registers + deterministic segmented reduction (no atomics).
//
// Replaces the reference's serial triple loop
//   Assembler<FEM,D,B,I>::discretize_operator   finite_elements/fem_assembler.h:52-121
//   Integrator::integrate_weak_form             utils/integration/integrator.h:93-106
//   Simplex::initialize (J, J^-1, measure)      geometry/simplex.h:184-195
//   weak forms                                  operators/{laplacian,diffusion,advection,reaction,dt}.h
//   Assembler::discretize_forcing               finite_elements/fem_assembler.h:122-136, integrator.h:74-90
//   FEMSolverBase::set_dirichlet_bc             solvers/fem_solver_base.h:144-155
// One thread owns one cell: vertex ids and coordinates are read coalesced from the struct-of-arrays copies,
// J / J^-1 / |det| live in registers, the reference-element tables (basis values, gradients, weights) are staged
// in shared memory, the quadrature loop is evaluated inline, and each local entry is written straight to its
// slot of the sorted contribution list (scatter map of pattern.cu).  A second kernel sums every segment left to
// right -- the order Eigen's setFromTriplets uses -- so repeated runs are bit-identical.
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
    const double* Kp;  // space-varying coefficient rows (device), row nq*e+q
    const double* bp;
    const double* cp;
};

template <int M> struct Geo {
    double invJ[M][M];  // invJ[m][r]
    double J[M][M];     // J[r][m]
    double x0[M];
    double measure;
};

template <int M>
__device__ __forceinline__ void load_geometry(int e, int n_cells, int n_nodes, const int32_t* __restrict__ verts,
                                              const double* __restrict__ coords, Geo<M>& g) {
    int v[M + 1];
#pragma unroll
    for (int k = 0; k <= M; ++k) v[k] = verts[(size_t)k * n_cells + e];
    double x[M + 1][M];
#pragma unroll
    for (int k = 0; k <= M; ++k)
#pragma unroll
        for (int r = 0; r < M; ++r) x[k][r] = __ldg(coords + (size_t)r * n_nodes + v[k]);
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

__device__ __forceinline__ void stage_tables(const FeTables* __restrict__ tab, FeTables* sm) {
    const int words = sizeof(FeTables) / sizeof(int);
    const int* src = reinterpret_cast<const int*>(tab);
    int* dst = reinterpret_cast<int*>(sm);
    for (int k = threadIdx.x; k < words; k += blockDim.x) dst[k] = src[k];
    __syncthreads();
}

constexpr __host__ __device__ int nbasis(int M, int R) { return R == 1 ? M + 1 : (M + 1) * (M + 2) / 2; }
constexpr __host__ __device__ int nquad(int M, int R) { return M == 2 ? (R == 1 ? 3 : 6) : (R == 1 ? 4 : 5); }

// ---- K3: one thread per cell, local matrix in registers ---------------------------------------------------------
template <int M, int R, bool SYM>
__global__ void __launch_bounds__(128)
k_local_assemble(int n_cells, int n_nodes, const int32_t* __restrict__ verts, const double* __restrict__ coords,
                 const FeTables* __restrict__ tab, OpCanon op, const int32_t* __restrict__ pos,
                 double* __restrict__ contrib) {
    constexpr int NB = nbasis(M, R), NQ = nquad(M, R);
    constexpr int NE = SYM ? NB * (NB + 1) / 2 : NB * NB;
    __shared__ FeTables T;
    stage_tables(tab, &T);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    Geo<M> geo;
    load_geometry<M>(e, n_cells, n_nodes, verts, coords, geo);

    double acc[NE];
#pragma unroll
    for (int s = 0; s < NE; ++s) acc[s] = 0.0;
    const bool need_grad = op.has_lap | op.has_diff | op.has_adv;

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
#pragma unroll
    for (int s = 0; s < NE; ++s) contrib[pos[(size_t)s * n_cells + e]] = acc[s] * geo.measure;
}

// ---- K3 for P2 tetrahedra (extension A10): 55/100 local entries do not fit in registers, so the physical
// gradients of all (q, i) are staged per thread in shared memory and the pair loop runs outermost.
template <bool SYM>
__global__ void __launch_bounds__(64)
k_local_assemble_p2tet(int n_cells, int n_nodes, const int32_t* __restrict__ verts, const double* __restrict__ coords,
                       const FeTables* __restrict__ tab, OpCanon op, const int32_t* __restrict__ pos,
                       double* __restrict__ contrib) {
    constexpr int M = 3, NB = 10, NQ = 5, TPB = 64;
    __shared__ FeTables T;
    extern __shared__ double dyn[];  // g[(q*NB+i)*3+r][tid], then kg, bg
    stage_tables(tab, &T);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    const int tid = threadIdx.x;
    Geo<M> geo;
    load_geometry<M>(e, n_cells, n_nodes, verts, coords, geo);
    double* G = dyn;                        // NQ*NB*3 * TPB
    double* KG = G + NQ * NB * 3 * TPB;     // NQ*NB*3 * TPB
    double* BG = KG + NQ * NB * 3 * TPB;    // NQ*NB * TPB
    for (int q = 0; q < NQ; ++q) {
        double K[9], bb[3];
        for (int k = 0; k < 9; ++k) K[k] = op.sv_diff ? op.Kp[((size_t)NQ * e + q) * 9 + k] : op.K[k];
        for (int r = 0; r < 3; ++r) bb[r] = op.sv_adv ? op.bp[((size_t)NQ * e + q) * 3 + r] : op.b[r];
        for (int i = 0; i < NB; ++i) {
            double gi[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double s = 0;
#pragma unroll
                for (int m = 0; m < 3; ++m) s += geo.invJ[m][r] * T.gref[(q * NB + i) * 3 + m];
                gi[r] = s;
                G[((q * NB + i) * 3 + r) * TPB + tid] = s;
            }
            if (op.has_diff) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    double s = 0;
#pragma unroll
                    for (int c = 0; c < 3; ++c) s += K[c * 3 + r] * gi[c];
                    KG[((q * NB + i) * 3 + r) * TPB + tid] = s;
                }
            }
            if (op.has_adv) {
                double s = 0;
#pragma unroll
                for (int r = 0; r < 3; ++r) s += gi[r] * bb[r];
                BG[(q * NB + i) * TPB + tid] = s;
            }
        }
    }
    int s_idx = 0;
    for (int i = 0; i < NB; ++i)
        for (int j = (SYM ? i : 0); j < NB; ++j) {
            double value = 0;
            for (int q = 0; q < NQ; ++q) {
                double val = 0.0;
                if (op.has_lap) {
                    double d = 0;
#pragma unroll
                    for (int r = 0; r < 3; ++r) d += G[((q * NB + i) * 3 + r) * TPB + tid] * G[((q * NB + j) * 3 + r) * TPB + tid];
                    val += op.s_lap * (-d);
                }
                if (op.has_diff) {
                    double d = 0;
#pragma unroll
                    for (int r = 0; r < 3; ++r) d += G[((q * NB + i) * 3 + r) * TPB + tid] * KG[((q * NB + j) * 3 + r) * TPB + tid];
                    val += op.s_diff * (-d);
                }
                if (op.has_adv) val += op.s_adv * (T.phi[q * NB + i] * BG[(q * NB + j) * TPB + tid]);
                if (op.has_reac) {
                    double cq = op.sv_reac ? op.cp[(size_t)NQ * e + q] : op.c;
                    val += op.s_reac * (cq * T.phi[q * NB + i] * T.phi[q * NB + j]);
                }
                value += val * T.w[q];
            }
            contrib[pos[(size_t)s_idx * n_cells + e]] = value * geo.measure;
            ++s_idx;
        }
}

// ---- K4: one thread per stored (unique) entry, left-to-right sum of its segment ---------------------------------
template <bool SYM>
__global__ void __launch_bounds__(256)
k_segmented_reduce(int64_t n_unique, const int32_t* __restrict__ seg, const double* __restrict__ contrib,
                   const int32_t* __restrict__ dst_a, const int32_t* __restrict__ dst_b, double* __restrict__ val) {
    int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (u >= n_unique) return;
    int t0 = seg[u], t1 = seg[u + 1];
    double s = contrib[t0];
    for (int t = t0 + 1; t < t1; ++t) s += contrib[t];
    val[dst_a[u]] = s;
    if constexpr (SYM) {
        int m = dst_b[u];
        if (m >= 0) val[m] = s;
    }
}

// ---- K5: load vector ---------------------------------------------------------------------------------------------
template <int M, int R>
__global__ void __launch_bounds__(128)
k_local_forcing(int n_cells, int n_nodes, const int32_t* __restrict__ verts, const double* __restrict__ coords,
                const FeTables* __restrict__ tab, const double* __restrict__ f_quad, const int32_t* __restrict__ pos,
                double* __restrict__ contrib) {
    constexpr int NB = nbasis(M, R), NQ = nquad(M, R);
    __shared__ FeTables T;
    stage_tables(tab, &T);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    Geo<M> geo;
    load_geometry<M>(e, n_cells, n_nodes, verts, coords, geo);
    double f[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) f[q] = f_quad[(size_t)NQ * e + q];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        double value = 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) value += (f[q] * T.phi[q * NB + i]) * T.w[q];
        contrib[pos[(size_t)i * n_cells + e]] = value * geo.measure;
    }
}

__global__ void __launch_bounds__(256)
k_reduce_forcing(int n_dofs, const int32_t* __restrict__ seg, const double* __restrict__ contrib,
                 double* __restrict__ b) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_dofs) return;
    double s = 0;
    for (int t = seg[d]; t < seg[d + 1]; ++t) s += contrib[t];
    b[d] = s;
}

// ---- quadrature nodes / dof coordinates --------------------------------------------------------------------------
template <int M, int R>
__global__ void k_quadrature_nodes(int n_cells, int n_nodes, const int32_t* __restrict__ verts,
                                   const double* __restrict__ coords, const FeTables* __restrict__ tab,
                                   double* __restrict__ out) {
    constexpr int NQ = nquad(M, R);
    __shared__ FeTables T;
    stage_tables(tab, &T);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    Geo<M> geo;
    load_geometry<M>(e, n_cells, n_nodes, verts, coords, geo);
    size_t rows = (size_t)n_cells * NQ;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int r = 0; r < M; ++r) {
            double s = 0;
#pragma unroll
            for (int m = 0; m < M; ++m) s += geo.J[r][m] * T.qn[q * M + m];
            out[(size_t)r * rows + (size_t)NQ * e + q] = s + geo.x0[r];
        }
}

__global__ void k_first_cell(int n_cells, int nb, int first_slot, const int32_t* __restrict__ dofs,
                             int32_t* __restrict__ first) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int ns = nb - first_slot;
    if (t >= (int64_t)n_cells * ns) return;
    int e = (int)(t / ns), j = first_slot + (int)(t % ns);
    atomicMin(&first[dofs[(size_t)j * n_cells + e]], e);  // min is order independent => deterministic
}

template <int M>
__global__ void k_edge_dof_coords(int n_cells, int n_nodes, int n_dofs, int nb, const int32_t* __restrict__ verts,
                                  const int32_t* __restrict__ dofs, const double* __restrict__ coords,
                                  const FeTables* __restrict__ tab, const int32_t* __restrict__ first,
                                  double* __restrict__ out) {
    __shared__ FeTables T;
    stage_tables(tab, &T);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    Geo<M> geo;
    load_geometry<M>(e, n_cells, n_nodes, verts, coords, geo);
    for (int j = M + 1; j < nb; ++j) {
        int d = dofs[(size_t)j * n_cells + e];
        if (first[d] != e) continue;
        for (int r = 0; r < M; ++r) {
            double s = 0;
            for (int m = 0; m < M; ++m) s += geo.J[r][m] * T.refn[j * M + m];
            out[(size_t)r * n_dofs + d] = s + geo.x0[r];
        }
    }
}

// ---- K6: Dirichlet rows --------------------------------------------------------------------------------------------
__global__ void k_dirichlet(int n, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                            const uint8_t* __restrict__ boundary, const double* __restrict__ g,
                            double* __restrict__ val, double* __restrict__ b, double* __restrict__ x0) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (!(r == 0 || boundary[r])) return;  // dof 0 is always visited (fem_solver_base.h:86)
    for (int t = rowptr[r]; t < rowptr[r + 1]; ++t) val[t] = (colidx[t] == r) ? 1.0 : 0.0;
    b[r] = g[r];
    if (x0) x0[r] = g[r];
}

// =====================================================================================================================
static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// lower the operator expression (fdb_opdesc) to its canonical form; uploads space-varying coefficient rows
static int canonicalize(fdb_space* s, const fdb_opdesc* d, OpCanon* o, std::vector<DevBuf<double>*>& keep) {
    memset(o, 0, sizeof(*o));
    const int N = s->N;
    FDB_CHECK(d && d->n_terms >= 0 && d->n_terms <= FDB_MAX_TERMS, FDB_ERR_ARG, "bad operator descriptor");
    const size_t qrows = (size_t)s->n_cells * s->nq;
    for (int t = 0; t < d->n_terms; ++t) {
        const fdb_term& T = d->terms[t];
        switch (T.kind) {
        case FDB_LAPLACIAN:
            o->s_lap = o->has_lap ? o->s_lap + T.scale : T.scale;
            o->has_lap = 1;
            break;
        case FDB_DIFFUSION:
        case FDB_ADVECTION:
        case FDB_REACTION: {
            FDB_CHECK(T.coeff != nullptr, FDB_ERR_ARG, "operator term without coefficient");
            int width = T.kind == FDB_DIFFUSION ? N * N : (T.kind == FDB_ADVECTION ? N : 1);
            int& has = T.kind == FDB_DIFFUSION ? o->has_diff : (T.kind == FDB_ADVECTION ? o->has_adv : o->has_reac);
            int& sv = T.kind == FDB_DIFFUSION ? o->sv_diff : (T.kind == FDB_ADVECTION ? o->sv_adv : o->sv_reac);
            double& sc = T.kind == FDB_DIFFUSION ? o->s_diff : (T.kind == FDB_ADVECTION ? o->s_adv : o->s_reac);
            double* cst = T.kind == FDB_DIFFUSION ? o->K : (T.kind == FDB_ADVECTION ? o->b : &o->c);
            if (T.space_varying) {
                FDB_CHECK(!has, FDB_ERR_UNSUPPORTED, "a space-varying term cannot be combined with another term of its kind");
                DevBuf<double>* buf = new DevBuf<double>();
                keep.push_back(buf);
                FDB_TRY(buf->alloc(qrows * width));
                FDB_CUDA(cudaMemcpyAsync(buf->p, T.coeff, sizeof(double) * qrows * width, cudaMemcpyHostToDevice, s->stream));
                (T.kind == FDB_DIFFUSION ? o->Kp : (T.kind == FDB_ADVECTION ? o->bp : o->cp)) = buf->p;
                sv = 1;
                sc = T.scale;
            } else if (!has) {
                for (int k = 0; k < width; ++k) cst[k] = T.coeff[k];
                sc = T.scale;
            } else {
                FDB_CHECK(!sv, FDB_ERR_UNSUPPORTED, "a space-varying term cannot be combined with another term of its kind");
                // fold: s1*c1 + s2*c2 with unit scale
                for (int k = 0; k < width; ++k) cst[k] = sc * cst[k] + T.scale * T.coeff[k];
                sc = 1.0;
            }
            has = 1;
        } break;
        case FDB_DT: break;  // zero field (operators/dt.h:34-36)
        default: FDB_CHECK(false, FDB_ERR_ARG, "unknown operator term kind");
        }
    }
    return FDB_OK;
}

template <int M, int R>
static int launch_local(fdb_space* s, const Pattern& P, const OpCanon& op, double* contrib) {
    const int B = 128;
    if (P.symmetric)
        k_local_assemble<M, R, true><<<grid_for(s->n_cells, B), B, 0, s->stream>>>(
            s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, op, P.pos.p, contrib);
    else
        k_local_assemble<M, R, false><<<grid_for(s->n_cells, B), B, 0, s->stream>>>(
            s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, op, P.pos.p, contrib);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

static int launch_local_p2tet(fdb_space* s, const Pattern& P, const OpCanon& op, double* contrib) {
    const int B = 64;
    const size_t dyn = sizeof(double) * B * (5 * 10 * 3 * 2 + 5 * 10);
    static bool configured = false;
    if (!configured) {
        FDB_CUDA(cudaFuncSetAttribute(k_local_assemble_p2tet<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        FDB_CUDA(cudaFuncSetAttribute(k_local_assemble_p2tet<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        configured = true;
    }
    if (P.symmetric)
        k_local_assemble_p2tet<true><<<grid_for(s->n_cells, B), B, dyn, s->stream>>>(
            s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, op, P.pos.p, contrib);
    else
        k_local_assemble_p2tet<false><<<grid_for(s->n_cells, B), B, dyn, s->stream>>>(
            s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, op, P.pos.p, contrib);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

static int ensure_contrib(fdb_space* s, size_t n) {
    if (s->contrib.n >= n) return FDB_OK;
    return s->contrib.alloc(n);
}

int assemble_operator(fdb_space* s, const fdb_opdesc* d, fdb_matrix* A) {
    FDB_CHECK(s && d && A && A->space == s, FDB_ERR_ARG, "fdb_assemble_operator: bad handle");
    const int sym = d->symmetric ? 1 : 0;
    FDB_TRY(build_pattern(s, sym));
    const Pattern& P = s->pat[sym];
    OpCanon op;
    std::vector<DevBuf<double>*> keep;
    int rc = canonicalize(s, d, &op, keep);
    if (rc == FDB_OK) rc = ensure_contrib(s, (size_t)P.n_contrib);
    if (rc == FDB_OK && (A->pat != &P || A->val.n < (size_t)P.nnz)) {
        rc = A->val.alloc((size_t)P.nnz);
        A->pat = &P;
    }
    if (rc == FDB_OK && s->profile) cudaEventRecord(s->ev[0], s->stream);
    if (rc == FDB_OK) {
        if (s->M == 2 && s->R == 1) rc = launch_local<2, 1>(s, P, op, s->contrib.p);
        else if (s->M == 2 && s->R == 2) rc = launch_local<2, 2>(s, P, op, s->contrib.p);
        else if (s->M == 3 && s->R == 1) rc = launch_local<3, 1>(s, P, op, s->contrib.p);
        else rc = launch_local_p2tet(s, P, op, s->contrib.p);
    }
    if (rc == FDB_OK && s->profile) cudaEventRecord(s->ev[1], s->stream);
    if (rc == FDB_OK) {
        const int B = 256;
        if (P.symmetric)
            k_segmented_reduce<true><<<grid_for(P.n_unique, B), B, 0, s->stream>>>(P.n_unique, P.seg.p, s->contrib.p,
                                                                                  P.dst_a.p, P.dst_b.p, A->val.p);
        else
            k_segmented_reduce<false><<<grid_for(P.n_unique, B), B, 0, s->stream>>>(P.n_unique, P.seg.p, s->contrib.p,
                                                                                   P.dst_a.p, nullptr, A->val.p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error(std::string("segmented reduce launch: ") + cudaGetErrorString(e)); rc = FDB_ERR_CUDA; }
    }
    if (rc == FDB_OK && s->profile) {
        cudaEventRecord(s->ev[2], s->stream);
        s->ev_valid = true;
    }
    if (!keep.empty()) {  // space-varying coefficient rows must outlive the kernels
        cudaStreamSynchronize(s->stream);
        for (auto* b : keep) delete b;
    }
    if (rc == FDB_OK) A->assembled = true;
    return rc;
}

int assemble_forcing(fdb_space* s, const double* f_quad, double* b) {
    FDB_TRY(build_forcing_map(s));
    const size_t total = (size_t)s->n_cells * s->nb;
    FDB_TRY(ensure_contrib(s, total));
    const int B = 128;
    const unsigned G = grid_for(s->n_cells, B);
#define FDB_LAUNCH_F(MM, RR)                                                                                       \
    k_local_forcing<MM, RR><<<G, B, 0, s->stream>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, f_quad, \
                                                    s->fmap.pos.p, s->contrib.p)
    if (s->M == 2 && s->R == 1) FDB_LAUNCH_F(2, 1);
    else if (s->M == 2 && s->R == 2) FDB_LAUNCH_F(2, 2);
    else if (s->M == 3 && s->R == 1) FDB_LAUNCH_F(3, 1);
    else FDB_LAUNCH_F(3, 2);
#undef FDB_LAUNCH_F
    FDB_CUDA(cudaGetLastError());
    k_reduce_forcing<<<grid_for(s->n_dofs, 256), 256, 0, s->stream>>>(s->n_dofs, s->fmap.seg.p, s->contrib.p, b);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

int quadrature_nodes(fdb_space* s, double* out) {
    const int B = 128;
    const unsigned G = grid_for(s->n_cells, B);
#define FDB_LAUNCH_Q(MM, RR) \
    k_quadrature_nodes<MM, RR><<<G, B, 0, s->stream>>>(s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, out)
    if (s->M == 2 && s->R == 1) FDB_LAUNCH_Q(2, 1);
    else if (s->M == 2 && s->R == 2) FDB_LAUNCH_Q(2, 2);
    else if (s->M == 3 && s->R == 1) FDB_LAUNCH_Q(3, 1);
    else FDB_LAUNCH_Q(3, 2);
#undef FDB_LAUNCH_Q
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

int dofs_coords(fdb_space* s, double* out) {
    // vertex dofs: the node coordinates themselves (lagrangian_basis.h:166)
    for (int r = 0; r < s->N; ++r)
        FDB_CUDA(cudaMemcpyAsync(out + (size_t)r * s->n_dofs, s->coords.p + (size_t)r * s->n_nodes,
                                 sizeof(double) * s->n_nodes, cudaMemcpyDeviceToDevice, s->stream));
    if (s->R == 1) return FDB_OK;
    DevBuf<int32_t> first;
    FDB_TRY(first.alloc(s->n_dofs));
    FDB_CUDA(cudaMemsetAsync(first.p, 0x7F, sizeof(int32_t) * s->n_dofs, s->stream));
    int ns = s->nb - (s->M + 1);
    k_first_cell<<<grid_for((int64_t)s->n_cells * ns, 256), 256, 0, s->stream>>>(s->n_cells, s->nb, s->M + 1, s->dofs.p,
                                                                                first.p);
    FDB_CUDA(cudaGetLastError());
    if (s->M == 2)
        k_edge_dof_coords<2><<<grid_for(s->n_cells, 128), 128, 0, s->stream>>>(
            s->n_cells, s->n_nodes, s->n_dofs, s->nb, s->verts_p, s->dofs.p, s->coords.p, s->tab.p, first.p, out);
    else
        k_edge_dof_coords<3><<<grid_for(s->n_cells, 128), 128, 0, s->stream>>>(
            s->n_cells, s->n_nodes, s->n_dofs, s->nb, s->verts_p, s->dofs.p, s->coords.p, s->tab.p, first.p, out);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    return FDB_OK;
}

int apply_dirichlet(fdb_matrix* A, const double* g, double* b, double* x0) {
    fdb_space* s = A->space;
    FDB_CHECK(A->assembled, FDB_ERR_STATE, "solver must be initialized first!");
    FDB_CHECK(s->has_boundary, FDB_ERR_STATE, "fdb_space_set_boundary has not been called");
    k_dirichlet<<<grid_for(s->n_dofs, 256), 256, 0, s->stream>>>(s->n_dofs, A->pat->rowptr.p, A->pat->colidx.p,
                                                                s->boundary.p, g, A->val.p, b, x0);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

}  // namespace fdb
