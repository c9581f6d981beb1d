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
#include "local_matrix.cuh"

namespace fdb {

// ---- K3 (two-kernel path): one thread per cell, local matrix in registers, written to the contribution list -------
// MODE_LEAN: closed form of the P1 stiffness matrix; MODE_TENSOR: constant coefficients, reference tensors staged in shared
// memory (entries are written as they are computed); MODE_QUAD: quadrature loop (space-varying coefficients).
template <int M, int R, bool SYM, int MODE>
__global__ void __launch_bounds__(128)
k_local_assemble(int n_cells, int n_nodes, const int32_t* __restrict__ verts, const double* __restrict__ coords,
                 const FeTables* __restrict__ tab, const double* __restrict__ tens, OpCanon op,
                 const int32_t* __restrict__ pos, double* __restrict__ contrib) {
    constexpr int NE = nentries(M, R, SYM), NB = nbasis(M, R);
    __shared__ FeTables T;
    constexpr int TSZ = is_tensor_mode(MODE) ? NB * NB * tens_stride_of(M, MODE) : 2;
    __shared__ __align__(16) double s_tens[TSZ];
    if constexpr (MODE == MODE_QUAD) stage_tables(tab, &T);
    if constexpr (is_tensor_mode(MODE)) stage_tensor_table(tens, s_tens, TSZ);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    double x[M + 1][M];
    gather_vertices<M>(e, n_cells, n_nodes, verts, coords, x);
    if constexpr (is_tensor_mode(MODE)) {
        Geo<M> geo;
        finish_geometry<M>(x, geo);
        TensWeights<M> w;
        tens_weights<M>(geo, op, w);
        int s_idx = 0;
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int j = (SYM ? i : 0); j < NB; ++j) {
                contrib[pos[(size_t)s_idx * n_cells + e]] = tens_entry<M, MODE>(s_tens, i * NB + j, w);
                ++s_idx;
            }
    } else {
        double acc[NE];
        cell_matrix<M, R, SYM, MODE == MODE_LEAN>(x, T, op, e, acc);
#pragma unroll
        for (int s = 0; s < NE; ++s) contrib[pos[(size_t)s * n_cells + e]] = acc[s];
    }
}

// ---- K3+K4 fused: one CTA per block of rows ----------------------------------------------------------------------
// Everything a CTA reads is one contiguous, block-major slice (built once with the pattern):
//   prologue: one thread hands the block's gather indices and segment offsets to the bulk-copy (TMA) engine, which
//             lands them in shared memory while phase 1 runs; completion is counted in bytes on an mbarrier;
//   phase 1 : the local matrices of every cell incident to the block's rows are computed into shared memory
//             (loc[slot * lcap + local cell], conflict free); vertex ids stream in, coordinates are gathered from
//             the packed (one 32-byte sector per node) copy.  Cells on the block boundary are recomputed by the
//             neighbouring block; nothing is exchanged through HBM;
//   phase 2 : one thread per stored entry of the block's rows sums its contributions from shared memory left to
//             right in emission order (ascending cell id) -- bit-identical to the two-kernel path and to Eigen's
//             setFromTriplets order -- and writes the value (and its mirror) once.
// bulk (TMA) copy global -> shared, completion counted in bytes on an mbarrier: the copy engine moves the block's
// lists, so they do not pass through the LSU pipe that limits this kernel
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@!p bra WAIT_%=;\n}" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// A block stores only the local entries it sums: every listed cell carries a mask of its needed emission slots and the
// start of its compact record in the block's shared-memory array (records of popcount | 1 doubles; cells are listed in
// descending mask order, so the lanes of a warp mostly share one mask: they skip the same entries and their records sit
// at one odd stride, i.e. conflict free).
// DSM: the block's destination list also rides the bulk-copy prologue into shared memory (used whenever the extra
// 4 / 8 bytes per entry do not cost a resident CTA); otherwise it is read from global memory.
// Destinations are (position, mirror position or -1) pairs for symmetric patterns, plain positions otherwise.
// Reference tensors in the constant bank: with the (i, j) loops unrolled every table value is an
// immediate constant-bank operand of its DFMA -- no shared-memory wavefronts for the tables.
constexpr int ct_off(int M, int R) { return (M == 2 && R == 1) ? 0 : (M == 2 && R == 2) ? 117 : (M == 3 && R == 1) ? 585 : 921; }
constexpr int CT_TOTAL = 3021;   // tens_total of (2,1) + (2,2) + (3,1) + (3,2)
__constant__ double c_tens[CT_TOTAL];

int upload_tensor_constants(int M, int R, const double* host, int count, cudaStream_t st) {
    FDB_CHECK(count == tens_total(M, nbasis(M, R)) && ct_off(M, R) + count <= CT_TOTAL, FDB_ERR_ARG, "tensor table size");
    FDB_CUDA(cudaMemcpyToSymbolAsync(c_tens, host, sizeof(double) * count, sizeof(double) * ct_off(M, R), cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    return FDB_OK;
}

template <int M, int R, int MODE>
__device__ __forceinline__ double tens_entry_c(int ij, const TensWeights<M>& w) {
    constexpr int OFF = ct_off(M, R) + tens_offset_of(M, nbasis(M, R), MODE);
    if constexpr (MODE == MODE_TENS_REAC) {
        return w.gamma * c_tens[OFF + ij];
    } else if constexpr (MODE == MODE_TENS_LAP) {
        constexpr int TS = tens_stride_of(M, MODE_TENS_LAP);
        const double* t = c_tens + OFF + ij * TS;
        if constexpr (M == 2) {
            double v = w.W[0] * t[0];
            v = fma(w.W[1], t[1], v);
            v = fma(w.W[3], t[2], v);
            return v;
        } else {
            double v = w.W[0] * t[0];
            v = fma(w.W[1], t[1], v);
            v = fma(w.W[2], t[2], v);
            v = fma(w.W[4], t[3], v);
            v = fma(w.W[5], t[4], v);
            v = fma(w.W[8], t[5], v);
            return v;
        }
    } else {
        constexpr int TS = tens_stride(M);
        const double* t = c_tens + OFF + ij * TS;
        double v = w.gamma * t[M * M + M];
#pragma unroll
        for (int n = 0; n < M; ++n) v += w.beta[n] * t[M * M + n];
#pragma unroll
        for (int k = 0; k < M * M; ++k) v += w.W[k] * t[k];
        return v;
    }
}
#ifdef FDB_TENS_SMEM   // development: stage the tables in shared memory instead (measured slower: C3 0.318 vs 0.283 ms)
constexpr bool TENS_CONST = false;
#else
constexpr bool TENS_CONST = true;
#endif

template <bool SYM> struct DstOf { using type = int2; };
template <> struct DstOf<false> { using type = int32_t; };

template <int M> struct CellRec { VertexIds<M> v; unsigned long long mask; int base; };
template <int M, bool COMPACT, bool NODES>
__device__ __forceinline__ CellRec<M> load_cell_rec(const int32_t* __restrict__ bverts, const uint16_t* __restrict__ bvloc,
                                                    const unsigned long long* __restrict__ bmask,
                                                    const uint16_t* __restrict__ bbase, size_t c) {
    CellRec<M> r;
    if constexpr (NODES) {   // four 16-bit block-local node indices: one 8-byte load
        const ushort4 q = __ldg(reinterpret_cast<const ushort4*>(bvloc) + c);
        r.v.v[0] = q.x; r.v.v[1] = q.y; r.v.v[2] = q.z;
        if constexpr (M == 3) r.v.v[3] = q.w;
    } else {
        r.v = load_vertex_ids<M>(bverts + c * (M + 1));
    }
    if constexpr (COMPACT) {
        r.mask = __ldg(bmask + c);
        r.base = (int)__ldg(bbase + c);
    } else {
        r.mask = 0; r.base = 0;
    }
    return r;
}

// coordinates of a cell's vertices from the block's node copies in shared memory (block-local indices)
// (x, y) pairs at s_xy[2 id], z at s_z[id] (s_z already carries the alignment shift of the block's copy)
template <int M>
__device__ __forceinline__ void gather_coords_shared(const VertexIds<M>& id, const double* s_xy, const double* s_z,
                                                     double (&x)[M + 1][M]) {
#pragma unroll
    for (int k = 0; k <= M; ++k) {
        const double2 a = *reinterpret_cast<const double2*>(s_xy + id.v[k] * 2);
        x[k][0] = a.x; x[k][1] = a.y;
        if constexpr (M == 3) x[k][2] = s_z[id.v[k]];
    }
}
// the two bulk copies that bring a block's node coordinates (first node n0, nnode nodes); returns the bytes requested
template <int M>
__device__ __forceinline__ unsigned request_coords(char* s_coords, int z_off, const double* __restrict__ bxy,
                                                   const double* __restrict__ bz, int n0, int nnode, uint64_t* bar) {
    const unsigned xyb = 16u * (unsigned)nnode;
    const unsigned zb = (M == 3) ? 16u * (unsigned)((nnode + (n0 & 1) + 1) >> 1) : 0u;
    if (nnode > 0) {
        bulk_copy_g2s(s_coords, bxy + (size_t)n0 * 2, xyb, bar);
        if constexpr (M == 3) bulk_copy_g2s(s_coords + z_off, bz + (size_t)(n0 & ~1), zb, bar);
    }
    return nnode > 0 ? xyb + zb : 0u;
}
template <int M>
__device__ __forceinline__ unsigned coords_bytes(int n0, int nnode) {
    if (nnode <= 0) return 0u;
    return 16u * (unsigned)nnode + ((M == 3) ? 16u * (unsigned)((nnode + (n0 & 1) + 1) >> 1) : 0u);
}

template <int M, int R, bool SYM, int MODE, bool DSM, int NTMAX, bool NODES>
// P1 triangles: <= 48 registers, 8 CTAs of 160 threads per SM; P2 tetrahedra: <= 64 registers, 4 CTAs of 256 threads;
// P1 tetrahedra (closed form): <= 56 registers, 3 CTAs of 384 threads (at 64 registers only 2 CTAs fit: 0.40 -> 0.49 ms on C4)
__global__ void __launch_bounds__(NTMAX, (M == 2 && R == 1 && MODE != MODE_QUAD) ? 5 : ((M == 3 && R == 2) ? 2 : ((M == 3 && R == 1 && MODE == MODE_LEAN) ? 3 : 1)))
k_fused_assemble(int ecap, int lcap, int con_cap, int ent_cap, const int32_t* __restrict__ bverts, const int32_t* __restrict__ bcells,
                 const unsigned long long* __restrict__ bmask, const uint16_t* __restrict__ bbase,
                 const double* __restrict__ coords_pk, const FeTables* __restrict__ tab, const double* __restrict__ tens,
                 OpCanon op, const int4* __restrict__ meta, const uint16_t* __restrict__ lidx,
                 const uint16_t* __restrict__ segrel, const typename DstOf<SYM>::type* __restrict__ dst,
                 double* __restrict__ val, const uint16_t* __restrict__ bvloc, const double* __restrict__ bcoords,
                 const double* __restrict__ bz, int coord_off, int z_off) {
    using Dst = typename DstOf<SYM>::type;
    constexpr int NE = nentries(M, R, SYM), NB = nbasis(M, R);
    constexpr int DPC = 16 / (int)sizeof(Dst);   // destinations per 16-byte chunk
    constexpr bool COMPACT = R == 2;             // P1: slot-major whole local matrices loc[slot * lcap + local cell]
    extern __shared__ double loc[];  // [ecap] compact records of the listed cells / [NE][lcap] local matrices
    uint16_t* s_lidx = reinterpret_cast<uint16_t*>(loc + ecap);
    uint16_t* s_seg = s_lidx + con_cap;
    Dst* s_dst = reinterpret_cast<Dst*>(s_seg + ent_cap);
    __shared__ FeTables T;
    const int b = blockIdx.x, NT = blockDim.x, tid = threadIdx.x;
    // one descriptor per block (two 16-byte loads): {first contribution, contributions, first entry, entries},
    // {first listed cell, listed cells, -, -}
    const int4 m0 = __ldg(meta + 2 * b), m1 = __ldg(meta + 2 * b + 1);
    const int c0 = m0.x, ncon = m0.y, e0 = m0.z, ne_b = m0.w, cc0 = m1.x, ncell = m1.y;
    // prologue: the block's gather indices, segment offsets (and destinations) go to shared memory through the bulk-copy
    // engine (UBLKCP issued by one thread, no LSU traffic); they are only needed in phase 2, so the copies overlap the
    // geometry gathers of phase 1
    const int base = c0 & ~7;                          // 16-byte aligned start of the 16-bit gather list
    const int dbase = e0 & ~(DPC - 1);                 // same for the destinations
    __shared__ uint64_t bar, bar2;
    const double* s_coords = reinterpret_cast<const double*>(reinterpret_cast<const char*>(loc) + coord_off);
    if (tid == 0) {
        if constexpr (NODES) {   // needed first: the node coordinates of the block (contiguous records per block)
            mbar_init(&bar2, 1);
            mbar_expect_tx(&bar2, coords_bytes<M>(m1.z, m1.w));
            request_coords<M>(reinterpret_cast<char*>(const_cast<double*>(s_coords)), z_off, bcoords, bz, m1.z, m1.w, &bar2);
        }
        const unsigned n16 = (unsigned)(c0 + ncon - base + 7) >> 3;   // 16-byte chunks
        const int sbase = (e0 + b) & ~7;               // same for the segment offsets (entries + 1 values)
        const unsigned s16 = (unsigned)(e0 + b + ne_b + 1 - sbase + 7) >> 3;
        const unsigned d16 = DSM ? (unsigned)(e0 + ne_b - dbase + DPC - 1) / DPC : 0u;
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, 16u * (n16 + s16 + d16));
        bulk_copy_g2s(s_lidx, lidx + base, 16u * n16, &bar);
        bulk_copy_g2s(s_seg, segrel + sbase, 16u * s16, &bar);
        if constexpr (DSM) bulk_copy_g2s(s_dst, dst + dbase, 16u * d16, &bar);
    }
    // reference-tensor form (constant coefficients): entries go straight to shared memory as they are computed
    constexpr bool CT = TENS_CONST;              // tables as constant-bank operands
    constexpr int TSZ = (is_tensor_mode(MODE) && !CT) ? NB * NB * tens_stride_of(M, MODE) : 2;
    __shared__ __align__(16) double s_tens[TSZ];
    if constexpr (is_tensor_mode(MODE) && !CT) stage_tensor_table(tens, s_tens, TSZ);
    if constexpr (MODE == MODE_QUAD) stage_tables(tab, &T);
    if constexpr (!DSM && R == 2) {   // the destinations are read in phase 2: pull their lines into L2 now
        const char* dp = reinterpret_cast<const char*>(dst + e0);
        const int nl = (ne_b * (int)sizeof(Dst) + 127) >> 7;
        for (int i = tid; i < nl; i += NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(dp + 128 * i));
    }
    // ---- phase 1: needed local entries of the block's cells -> shared memory -----------------------------------------
    // the record of a thread's next cell is requested before the coordinates of the current one are waited for
    // (splitting the cells that need most of their entries between two threads was measured slower: the second gather of
    // the coordinates costs more than the shorter critical path of phase 1 gains -- C3 0.318 vs 0.279 ms)
    CellRec<M> nxt;
    if (tid < ncell) nxt = load_cell_rec<M, COMPACT, NODES>(bverts, bvloc, bmask, bbase, (size_t)cc0 + tid);
    if constexpr (NODES) {
        __syncthreads();          // bar2 is initialised
        mbar_wait(&bar2, 0);      // the block's node coordinates have landed
    }
    for (int lc = tid; lc < ncell; lc += NT) {
        double x[M + 1][M];
        const CellRec<M> cur = nxt;
        if (lc + NT < ncell) nxt = load_cell_rec<M, COMPACT, NODES>(bverts, bvloc, bmask, bbase, (size_t)cc0 + lc + NT);
        if constexpr (NODES) gather_coords_shared<M>(cur.v, s_coords, s_coords + (z_off >> 3) + (m1.z & 1), x);
        else gather_coords_packed<M>(cur.v, coords_pk, x);
        double* rec = loc + (COMPACT ? cur.base : lc);
        if constexpr (is_tensor_mode(MODE)) {
            Geo<M> geo;
            finish_geometry<M>(x, geo);
            TensWeights<M> w;
            tens_weights<M>(geo, op, w);
            int s_idx = 0;
#pragma unroll
            for (int i = 0; i < NB; ++i)
#pragma unroll
                for (int j = (SYM ? i : 0); j < NB; ++j) {
                    if constexpr (COMPACT) {
                        if ((cur.mask >> s_idx) & 1ull) {
                            if constexpr (CT) *rec++ = tens_entry_c<M, R, MODE>(i * NB + j, w);
                            else *rec++ = tens_entry<M, MODE>(s_tens, i * NB + j, w);
                        }
                    } else {
                        if constexpr (CT) rec[s_idx * lcap] = tens_entry_c<M, R, MODE>(i * NB + j, w);
                        else rec[s_idx * lcap] = tens_entry<M, MODE>(s_tens, i * NB + j, w);
                    }
                    ++s_idx;
                }
        } else {
            int e = 0;
            if constexpr (MODE == MODE_QUAD) e = __ldg(bcells + cc0 + lc);
            double acc[NE];
            cell_matrix<M, R, SYM, MODE == MODE_LEAN>(x, T, op, e, acc);
#pragma unroll
            for (int s = 0; s < NE; ++s) {
                if constexpr (COMPACT) {
                    if ((cur.mask >> s) & 1ull) *rec++ = acc[s];
                } else {
                    rec[s * lcap] = acc[s];
                }
            }
        }
    }
    Dst dn{};
    if constexpr (!DSM && COMPACT) { if (tid < ne_b) dn = __ldg(dst + e0 + tid); }
    __syncthreads();
    mbar_wait(&bar, 0);
    // ---- phase 2: one thread per stored entry ---------------------------------------------------------------------------
    const int shift = c0 - base, sshift = (e0 + b) & 7, dshift = e0 - dbase;
    for (int k = tid; k < ne_b; k += NT) {
        int t = s_seg[k + sshift] + shift;
        const int t1 = s_seg[k + sshift + 1] + shift;
        Dst d;
        if constexpr (DSM) d = s_dst[k + dshift];
        else if constexpr (COMPACT) { d = dn; if (k + NT < ne_b) dn = __ldg(dst + e0 + k + NT); }   // one entry ahead (L2 hits)
        else d = __ldg(dst + e0 + k);
        double sum = loc[s_lidx[t]];
        for (++t; t < t1; ++t) sum += loc[s_lidx[t]];
        if constexpr (SYM) {
            val[d.x] = sum;
            if (d.y >= 0) val[d.y] = sum;
        } else {
            val[d] = sum;
        }
    }
}

// ---- persistent form of the fused kernel (constant coefficients: P1 slot-major layout and P2 compact records) --------
// One CTA per SM slot walks the row blocks b = blockIdx.x, blockIdx.x + gridDim.x, ...  Everything a block reads -- the
// block-local node indices of its cells (P2: also their slot masks and record starts), the coordinates of its nodes, the
// gather / segment / destination lists -- is a
// contiguous record in HBM and reaches shared memory through the bulk-copy engine, one block AHEAD of its use: the
// phase-1 inputs of block b+1 are requested when phase 1 of block b has finished (they land during phase 2), the phase-2
// inputs of block b+1 when phase 2 of block b has finished (they land during the next phase 1).  No thread waits on a
// global load, so the long-scoreboard stalls of the gathers (40 % of the samples of the plain fused kernel) disappear and
// single buffers suffice.  Arithmetic and summation order are those of k_fused_assemble: bit-identical results.
template <int M, int R, bool SYM, int MODE, int NT>
__global__ void __launch_bounds__(NT, NT > 512 ? 1 : (NT <= 256 ? 4 : 2))
k_fused_persist(int nblocks, PersistLayout L, const uint16_t* __restrict__ bvloc, const unsigned long long* __restrict__ bmask,
                const uint16_t* __restrict__ bbase, const double* __restrict__ bcoords, const double* __restrict__ bz,
                OpCanon op, const int4* __restrict__ meta, const uint16_t* __restrict__ lidx,
                const uint16_t* __restrict__ segrel, const typename DstOf<SYM>::type* __restrict__ dst,
                double* __restrict__ val) {
    using Dst = typename DstOf<SYM>::type;
    constexpr int NE = nentries(M, R, SYM), NB = nbasis(M, R);
    constexpr int DPC = 16 / (int)sizeof(Dst);
    constexpr bool COMPACT = R == 2;   // P2: compact records of the needed entries (slot masks), P1: slot-major local matrices
    extern __shared__ double loc[];
    char* sm = reinterpret_cast<char*>(loc);
    const uint16_t* s_ids = reinterpret_cast<const uint16_t*>(sm + L.off_ids);
    const unsigned long long* s_mask = reinterpret_cast<const unsigned long long*>(sm + L.off_mask);
    const uint16_t* s_base = reinterpret_cast<const uint16_t*>(sm + L.off_base);
    const double* s_coords = reinterpret_cast<const double*>(sm + L.off_coords);
    const uint16_t* s_lidx = reinterpret_cast<const uint16_t*>(sm + L.off_lidx);
    const uint16_t* s_seg = reinterpret_cast<const uint16_t*>(sm + L.off_seg);
    const Dst* s_dst = reinterpret_cast<const Dst*>(sm + L.off_dst);
    __shared__ uint64_t barA, barB;
    __shared__ int4 s_meta[2][2];
    const int tid = threadIdx.x, lcap = L.lcap_cells;
    int b = blockIdx.x;
    if (b >= nblocks) return;
    // phase-1 inputs of a block: node indices of its cells (8 bytes per cell), P2: slot masks and record starts,
    // coordinates of its nodes
    auto request_a = [&](const int4& m1) {
        const int cc0 = m1.x, ncell = m1.y;
        const unsigned ib = 16u * (unsigned)((ncell + (cc0 & 1) + 1) >> 1);
        const unsigned bb = COMPACT ? 16u * (unsigned)((ncell + (cc0 & 7) + 7) >> 3) : 0u;
        mbar_expect_tx(&barA, ib + (COMPACT ? ib + bb : 0u) + coords_bytes<M>(m1.z, m1.w));
        if (ib) bulk_copy_g2s(sm + L.off_ids, bvloc + (size_t)(cc0 & ~1) * 4, ib, &barA);   // (a block of rows without cells
        if constexpr (COMPACT) {                                                             //  requests nothing: the
            if (ib) bulk_copy_g2s(sm + L.off_mask, bmask + (size_t)(cc0 & ~1), ib, &barA);  //  expect_tx of 0 bytes alone
            if (bb) bulk_copy_g2s(sm + L.off_base, bbase + (size_t)(cc0 & ~7), bb, &barA);  //  completes the phase)
        }
        request_coords<M>(sm + L.off_coords, L.z_off, bcoords, bz, m1.z, m1.w, &barA);
    };
    // phase-2 inputs: gather indices, segment offsets, destinations
    auto request_b = [&](int blk, const int4& m0) {
        const int c0 = m0.x, ncon = m0.y, e0 = m0.z, ne_b = m0.w;
        const int base = c0 & ~7, sbase = (e0 + blk) & ~7, dbase = e0 & ~(DPC - 1);
        const unsigned n16 = (unsigned)(c0 + ncon - base + 7) >> 3;
        const unsigned s16 = (unsigned)(e0 + blk + ne_b + 1 - sbase + 7) >> 3;
        const unsigned d16 = (unsigned)(e0 + ne_b - dbase + DPC - 1) / DPC;
        mbar_expect_tx(&barB, 16u * (n16 + s16 + d16));
        if (n16) bulk_copy_g2s(sm + L.off_lidx, lidx + base, 16u * n16, &barB);
        bulk_copy_g2s(sm + L.off_seg, segrel + sbase, 16u * s16, &barB);
        if (d16) bulk_copy_g2s(sm + L.off_dst, dst + dbase, 16u * d16, &barB);
    };
    if (tid == 0) {
        mbar_init(&barA, 1);
        mbar_init(&barB, 1);
        const int4 m0 = __ldg(meta + 2 * b), m1 = __ldg(meta + 2 * b + 1);
        s_meta[0][0] = m0; s_meta[0][1] = m1;
        request_a(m1);
        request_b(b, m0);
    }
    __syncthreads();
    for (int it = 0; b < nblocks; ++it, b += gridDim.x) {
        const int cur = it & 1;
        const unsigned parity = (unsigned)(it & 1);
        const int4 m0 = s_meta[cur][0], m1 = s_meta[cur][1];
        const int c0 = m0.x, e0 = m0.z, ne_b = m0.w, cc0 = m1.x, ncell = m1.y;
        const int nb = b + (int)gridDim.x;
        int4 n0{}, n1{};
        if (tid == 0 && nb < nblocks) { n0 = __ldg(meta + 2 * nb); n1 = __ldg(meta + 2 * nb + 1); }   // used after phase 1
        mbar_wait(&barA, parity);
        // ---- phase 1 ----
        const int ishift = cc0 & 1, bshift = cc0 & 7;
        const double* s_z = s_coords + (L.z_off >> 3) + (m1.z & 1);
        for (int lc = tid; lc < ncell; lc += NT) {
            double x[M + 1][M];
            VertexIds<M> id;
            {
                const ushort4 q = *reinterpret_cast<const ushort4*>(s_ids + (size_t)(lc + ishift) * 4);
                id.v[0] = q.x; id.v[1] = q.y; id.v[2] = q.z;
                if constexpr (M == 3) id.v[3] = q.w;
            }
            gather_coords_shared<M>(id, s_coords, s_z, x);
            if constexpr (COMPACT) {
                const unsigned long long mask = s_mask[lc + ishift];
                double* rec = loc + s_base[lc + bshift];
                Geo<M> geo;
                finish_geometry<M>(x, geo);
                TensWeights<M> w;
                tens_weights<M>(geo, op, w);
                int s_idx = 0;
#pragma unroll
                for (int i = 0; i < NB; ++i)
#pragma unroll
                    for (int j = (SYM ? i : 0); j < NB; ++j) {
                        if ((mask >> s_idx) & 1ull) *rec++ = tens_entry_c<M, R, MODE>(i * NB + j, w);
                        ++s_idx;
                    }
            } else if constexpr (is_tensor_mode(MODE)) {
                double* rec = loc + lc;
                Geo<M> geo;
                finish_geometry<M>(x, geo);
                TensWeights<M> w;
                tens_weights<M>(geo, op, w);
                int s_idx = 0;
#pragma unroll
                for (int i = 0; i < NB; ++i)
#pragma unroll
                    for (int j = (SYM ? i : 0); j < NB; ++j) {
                        rec[s_idx * lcap] = tens_entry_c<M, R, MODE>(i * NB + j, w);
                        ++s_idx;
                    }
            } else {
                double* rec = loc + lc;
                double acc[NE];
                p1_laplacian_matrix<M, SYM>(x, op.lap_k0, acc);   // closed form of the P1 stiffness (MODE_LEAN)
#pragma unroll
                for (int s2 = 0; s2 < NE; ++s2) rec[s2 * lcap] = acc[s2];
            }
        }
        __syncthreads();   // local matrices complete; the phase-1 inputs are free
        if (tid == 0 && nb < nblocks) {
            s_meta[cur ^ 1][0] = n0; s_meta[cur ^ 1][1] = n1;
            request_a(n1);
        }
        mbar_wait(&barB, parity);
        // ---- phase 2 ----
        const int shift = c0 - (c0 & ~7), sshift = (e0 + b) & 7, dshift = e0 - (e0 & ~(DPC - 1));
        for (int k = tid; k < ne_b; k += NT) {
            int t = s_seg[k + sshift] + shift;
            const int t1 = s_seg[k + sshift + 1] + shift;
            const Dst d = s_dst[k + dshift];
            double sum = loc[s_lidx[t]];
            // (four index / value loads in flight per thread were measured slower: C4 0.3055 vs 0.3004 ms, C3 0.250 vs 0.241)
            ++t;
            for (; t < t1; ++t) sum += loc[s_lidx[t]];
            if constexpr (SYM) {
                val[d.x] = sum;
                if (d.y >= 0) val[d.y] = sum;
            } else {
                val[d] = sum;
            }
        }
        __syncthreads();   // local matrices and phase-2 inputs are free; s_meta[cur ^ 1] is visible
        if (tid == 0 && nb < nblocks) request_b(nb, s_meta[cur ^ 1][0]);
    }
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

// ---- K3 for P2 tetrahedra, constant coefficients (contribution-list path) ------------------------------------------
template <bool SYM, int MODE>
__global__ void __launch_bounds__(256)
k_local_assemble_p2tet_const(int n_cells, int n_nodes, const int32_t* __restrict__ verts, const double* __restrict__ coords,
                             const double* __restrict__ tens, OpCanon op, const int32_t* __restrict__ pos,
                             double* __restrict__ contrib) {
    constexpr int NB = 10, TSZ = NB * NB * tens_stride_of(3, MODE);
    __shared__ __align__(16) double tab[TSZ];
    stage_tensor_table(tens, tab, TSZ);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    Geo<3> geo;
    load_geometry<3>(e, n_cells, n_nodes, verts, coords, geo);
    TensWeights<3> w;
    tens_weights<3>(geo, op, w);
    int s_idx = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i)
#pragma unroll
        for (int j = (SYM ? i : 0); j < NB; ++j) {
            contrib[pos[(size_t)s_idx * n_cells + e]] = tens_entry<3, MODE>(tab, i * NB + j, w);
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

// first_slot = M + 1: edge dofs only (reference numbering, vertex dofs are the nodes); first_slot = 0: every dof of a
// renumbered table, vertex dofs copied from their node (exact), edge dofs mapped from the reference element
template <int M>
__global__ void k_edge_dof_coords(int n_cells, int n_nodes, int n_dofs, int nb, int first_slot, const int32_t* __restrict__ verts,
                                  const int32_t* __restrict__ dofs, const double* __restrict__ coords,
                                  const FeTables* __restrict__ tab, const int32_t* __restrict__ first,
                                  double* __restrict__ out) {
    __shared__ FeTables T;
    stage_tables(tab, &T);
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_cells) return;
    Geo<M> geo;
    load_geometry<M>(e, n_cells, n_nodes, verts, coords, geo);
    for (int j = first_slot; j < nb; ++j) {
        int d = dofs[(size_t)j * n_cells + e];
        if (first[d] != e) continue;
        if (j <= M) {
            const int v = verts[(size_t)j * n_cells + e];
            for (int r = 0; r < M; ++r) out[(size_t)r * n_dofs + d] = coords[(size_t)r * n_nodes + v];
            continue;
        }
        for (int r = 0; r < M; ++r) {
            double s = 0;
            for (int m = 0; m < M; ++m) s += geo.J[r][m] * T.refn[j * M + m];
            out[(size_t)r * n_dofs + d] = s + geo.x0[r];
        }
    }
}

// ---- K6: Dirichlet rows --------------------------------------------------------------------------------------------
__global__ void k_dirichlet(int n, int dof0_rule, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                            const uint8_t* __restrict__ boundary, const double* __restrict__ g,
                            double* __restrict__ val, double* __restrict__ b, double* __restrict__ x0) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (!((r == 0 && dof0_rule) || boundary[r])) return;  // dof 0 is always visited (fem_solver_base.h:86)
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
    for (int q = 0; q < s->nq; ++q) o->wsum += s->tab_host.w[q];  // left to right, like the quadrature loop
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
    o->lap_k0 = o->s_lap * o->wsum / (s->M == 3 ? 6.0 : 2.0);
    return FDB_OK;
}

template <int M, int R, bool SYM, int MODE>
static int launch_two_kernel_local(fdb_space* s, const Pattern& P, const OpCanon& op, double* contrib) {
    if constexpr (M == 3 && R == 2) {   // P2 tetrahedra: one thread per cell, entries written as they are computed
        if constexpr (is_tensor_mode(MODE)) {
            const int Bc = 256;
            k_local_assemble_p2tet_const<SYM, MODE><<<grid_for(s->n_cells, Bc), Bc, 0, s->stream>>>(
                s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tens.p + tens_offset_of(3, 10, MODE), op, P.pos.p, contrib);
        } else {
            const int B = 64;
            const size_t dyn = sizeof(double) * B * (5 * 10 * 3 * 2 + 5 * 10);
            static bool configured = false;
            if (!configured) {
                FDB_CUDA(cudaFuncSetAttribute(k_local_assemble_p2tet<SYM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                configured = true;
            }
            k_local_assemble_p2tet<SYM><<<grid_for(s->n_cells, B), B, dyn, s->stream>>>(
                s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, op, P.pos.p, contrib);
        }
    } else {
        const int B = 128;
        k_local_assemble<M, R, SYM, MODE><<<grid_for(s->n_cells, B), B, 0, s->stream>>>(
            s->n_cells, s->n_nodes, s->verts_p, s->coords.p, s->tab.p, s->tens.p + tens_offset_of(M, nbasis(M, R), MODE), op,
            P.pos.p, contrib);
    }
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

template <int M, int R, bool SYM, int MODE, bool DSM, bool NODES = false>
static int launch_fused_dsm(fdb_space* s, const Pattern& P, const OpCanon& op, double* val) {
    if constexpr (!NODES && R == 1) {
        if (P.f_nodes) return launch_fused_dsm<M, R, SYM, MODE, DSM, true>(s, P, op, val);
    }
    constexpr int NTMAX = (MODE == MODE_LEAN && M == 3) ? 384 : ((M == 3 && R == 2) ? 512 : 256);
    int con_cap, ent_cap;
    const size_t dyn = fused_smem_bytes(P, DSM, &con_cap, &ent_cap);
    static size_t configured = 0;
    if (dyn > configured) {
        FDB_CUDA(cudaFuncSetAttribute(k_fused_assemble<M, R, SYM, MODE, DSM, NTMAX, NODES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        configured = dyn;
    }
    int nt = s->fused_threads > 0 ? s->fused_threads : P.f_threads;
    if (nt > NTMAX) nt = NTMAX;
    using Dst = typename DstOf<SYM>::type;
    const Dst* dst;
    if constexpr (SYM) dst = P.f_dst.p; else dst = P.f_dst1.p;
    k_fused_assemble<M, R, SYM, MODE, DSM, NTMAX, NODES><<<P.f_nblocks, nt, dyn, s->stream>>>(
        P.f_lcap, P.f_cells_cap, con_cap, ent_cap, P.f_bverts.p, P.f_bcells.p, P.f_bmask.p, P.f_bbase.p, s->coords_pk.p, s->tab.p,
        s->tens.p + tens_offset_of(M, nbasis(M, R), MODE), op, reinterpret_cast<const int4*>(P.f_meta.p), P.f_lidx.p,
        P.f_segrel.p, dst, val, P.f_bvloc.p, P.f_bcoords.p, P.f_bz.p, (int)fused_coord_offset(P, DSM), P.f_node_z_off);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

// *handled = false: the block lists leave room for fewer than two CTAs per SM (the plain fused kernel takes the plan)
template <int M, int R, bool SYM, int MODE, int NT = (R == 1 ? 320 : 256)>
static int launch_fused_persist(fdb_space* s, const Pattern& P, const OpCanon& op, double* val, bool* handled) {
    *handled = false;
    if constexpr (NT == (R == 1 ? 320 : 256)) {   // development: other CTA sizes (C4 at 80 rows: 320 0.301, 384 0.308, 448 0.310 ms)
        static const int want = getenv("FDB_PERSIST_NT") ? atoi(getenv("FDB_PERSIST_NT")) : NT;
        if (want == 384 && NT != 384) return launch_fused_persist<M, R, SYM, MODE, 384>(s, P, op, val, handled);
        if (want == 512 && NT != 512) return launch_fused_persist<M, R, SYM, MODE, 512>(s, P, op, val, handled);
    }
    using Dst = typename DstOf<SYM>::type;
    PersistLayout L;
    const size_t dyn = persist_layout(P, &L);
    if (2 * (dyn + 1024) > 228 * 1024) return FDB_OK;
    static size_t configured = 0;
    static int per_sm = 0;
    if (dyn > configured) {
        FDB_CUDA(cudaFuncSetAttribute(k_fused_persist<M, R, SYM, MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        FDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused_persist<M, R, SYM, MODE, NT>, NT, dyn));
        configured = dyn;
        if (getenv("FDB_VERBOSE")) fprintf(stderr, "[fdb] persistent fused kernel: %zu B of shared memory, %d CTAs of %d threads per SM\n", dyn, per_sm, NT);
    }
    if (per_sm < 2) return FDB_OK;
    *handled = true;
    s->last_persist = true;
    int grid = per_sm * s->sm_count;
    if (grid > P.f_nblocks) grid = P.f_nblocks;
    const Dst* dst;
    if constexpr (SYM) dst = P.f_dst.p; else dst = P.f_dst1.p;
    k_fused_persist<M, R, SYM, MODE, NT><<<grid, NT, dyn, s->stream>>>(P.f_nblocks, L, P.f_bvloc.p, P.f_bmask.p, P.f_bbase.p,
        P.f_bcoords.p, P.f_bz.p, op,
        reinterpret_cast<const int4*>(P.f_meta.p), P.f_lidx.p, P.f_segrel.p, dst, val);
    FDB_CUDA(cudaGetLastError());
    return FDB_OK;
}

template <int M, int R, bool SYM, int MODE>
static int launch_fused(fdb_space* s, const Pattern& P, const OpCanon& op, double* val) {
    // P1 elements: the persistent prefetching kernel for the stiffness closed form, the mass matrix and the general
    // reference-tensor rows (tetrahedra, advection-diffusion-reaction: 0.96 vs 1.50 ms on the C4 mesh; two CTAs of 320
    // threads leave 102 registers per thread; triangles: C2 0.080 vs 0.097 ms); a pure diffusion tensor stays on the plain
    // kernel (0.63 vs 0.68 ms on tetrahedra).  FDB_FUSED_PERSIST=0: plain kernel everywhere.
    if constexpr (R == 1 && MODE != MODE_QUAD && MODE != MODE_TENS_LAP) {
        static const int persist = getenv("FDB_FUSED_PERSIST") ? atoi(getenv("FDB_FUSED_PERSIST")) : 1;
        if (P.f_nodes && persist != 0) {
            bool handled = false;
            FDB_TRY((launch_fused_persist<M, R, SYM, MODE>(s, P, op, val, &handled)));
            if (handled) return FDB_OK;
        }
    }
    // P2 elements, constant coefficients (opt-in while being measured: FDB_FUSED_PERSIST_P2=1 builds the node copies)
    if constexpr (R == 2 && is_tensor_mode(MODE) && nentries(M, R, SYM) <= 57) {
        if (P.f_nodes) {
            bool handled = false;
            FDB_TRY((launch_fused_persist<M, R, SYM, MODE>(s, P, op, val, &handled)));
            if (handled) return FDB_OK;
        }
    }
    if constexpr ((M == 3 && R == 2 && !is_tensor_mode(MODE)) || nentries(M, R, SYM) > 64) {
        set_error("no fused assembly for this space / operator (P2 tetrahedra: symmetric, constant coefficients)");
        return FDB_ERR_UNSUPPORTED;
    } else {
        if (P.f_dsm) return launch_fused_dsm<M, R, SYM, MODE, true>(s, P, op, val);
        return launch_fused_dsm<M, R, SYM, MODE, false>(s, P, op, val);
    }
}

// dispatch on (M, R, symmetric, evaluation mode): the closed form for the P1 stiffness matrix, the reference-tensor forms
// for every other constant-coefficient operator, the quadrature loop when a coefficient varies in space
#define FDB_DISPATCH_SYM(FN, MM, RR, MODE_, ...)                                                    \
    do {                                                                                            \
        if (P.symmetric) return FN<MM, RR, true, MODE_>(__VA_ARGS__);                               \
        return FN<MM, RR, false, MODE_>(__VA_ARGS__);                                               \
    } while (0)
#define FDB_DISPATCH_MR(FN, MM, RR, ...)                                                            \
    do {                                                                                            \
        if (mode == MODE_LEAN) {                                                                    \
            if constexpr (RR == 1) FDB_DISPATCH_SYM(FN, MM, RR, MODE_LEAN, __VA_ARGS__);            \
        } else if (mode == MODE_TENSOR) FDB_DISPATCH_SYM(FN, MM, RR, MODE_TENSOR, __VA_ARGS__);     \
        else if (mode == MODE_TENS_LAP) FDB_DISPATCH_SYM(FN, MM, RR, MODE_TENS_LAP, __VA_ARGS__);   \
        else if (mode == MODE_TENS_REAC) FDB_DISPATCH_SYM(FN, MM, RR, MODE_TENS_REAC, __VA_ARGS__); \
        else FDB_DISPATCH_SYM(FN, MM, RR, MODE_QUAD, __VA_ARGS__);                                  \
    } while (0)
#define FDB_DISPATCH(FN, ...)                                                                       \
    do {                                                                                            \
        if (s->M == 2 && s->R == 1) FDB_DISPATCH_MR(FN, 2, 1, __VA_ARGS__);                         \
        else if (s->M == 2 && s->R == 2) FDB_DISPATCH_MR(FN, 2, 2, __VA_ARGS__);                    \
        else if (s->M == 3 && s->R == 1) FDB_DISPATCH_MR(FN, 3, 1, __VA_ARGS__);                    \
        else if (s->M == 3 && s->R == 2) FDB_DISPATCH_MR(FN, 3, 2, __VA_ARGS__);                    \
    } while (0)

static int run_two_kernel_local(fdb_space* s, const Pattern& P, const OpCanon& op, int mode, double* contrib) {
    FDB_DISPATCH(launch_two_kernel_local, s, P, op, contrib);
    FDB_CHECK(false, FDB_ERR_UNSUPPORTED, "unsupported (M, R)");
}
static int run_fused(fdb_space* s, const Pattern& P, const OpCanon& op, int mode, double* val) {
    FDB_DISPATCH(launch_fused, s, P, op, val);
    FDB_CHECK(false, FDB_ERR_UNSUPPORTED, "unsupported (M, R)");
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
    if (rc == FDB_OK && (A->pat != &P || A->val.n < (size_t)P.nnz)) {
        rc = A->val.alloc((size_t)P.nnz);
        A->pat = &P;
    }
    const bool lap_only = op.has_lap && !op.has_diff && !op.has_adv && !op.has_reac;
    const bool varying = op.sv_diff || op.sv_adv || op.sv_reac;
    const bool reac_only = op.has_reac && !op.has_lap && !op.has_diff && !op.has_adv;
    const int mode = (lap_only && s->R == 1) ? MODE_LEAN
                     : varying ? MODE_QUAD
                     : lap_only ? MODE_TENS_LAP
                     : reac_only ? MODE_TENS_REAC : MODE_TENSOR;
    const bool p2tet = (s->M == 3 && s->R == 2);
    const bool surface = s->N != s->M;   // manifold cells: contribution-list path with the kernels of surface.cu
    // P2 tetrahedra with space-varying coefficients keep the contribution-list path (staged quadrature kernel)
    const bool no_fuse = surface || s->force_two_kernel || (p2tet && varying);
    Pattern& Pm = s->pat[sym];
    if (rc == FDB_OK && !no_fuse && Pm.n_assemblies >= 1) rc = ensure_fused_plan(s, &Pm);
    ++Pm.n_assemblies;
    const bool fused = P.fused && !no_fuse;
    if (rc == FDB_OK && !fused) rc = ensure_contrib(s, (size_t)P.n_contrib);
    if (rc == FDB_OK && s->profile) cudaEventRecord(s->ev[0], s->stream);
    s->last_persist = false;
    if (rc == FDB_OK) {
        if (fused) rc = run_fused(s, P, op, mode, A->val.p);
        else if (surface) rc = surface_local_assemble(s, P, op, s->contrib.p);
        else rc = run_two_kernel_local(s, P, op, mode, s->contrib.p);
    }
    if (rc == FDB_OK && s->profile) cudaEventRecord(s->ev[1], s->stream);
    if (rc == FDB_OK && !fused) {
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
    if (rc == FDB_OK) { A->assembled = true; ++A->val_version; s->last_fused = fused ? (s->last_persist ? 2 : 1) : 0; s->last_launches = fused ? 1 : 2; }
    return rc;
}

int assemble_forcing(fdb_space* s, const double* f_quad, double* b) {
    FDB_TRY(build_forcing_map(s));
    const size_t total = (size_t)s->n_cells * s->nb;
    FDB_TRY(ensure_contrib(s, total));
    const int B = 128;
    const unsigned G = grid_for(s->n_cells, B);
    if (s->N != s->M) {
        FDB_TRY(surface_local_forcing(s, f_quad, s->fmap.pos.p, s->contrib.p));
        k_reduce_forcing<<<grid_for(s->n_dofs, 256), 256, 0, s->stream>>>(s->n_dofs, s->fmap.seg.p, s->contrib.p, b);
        FDB_CUDA(cudaGetLastError());
        return FDB_OK;
    }
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
    if (s->N != s->M) return surface_quadrature_nodes(s, out);
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
    // reference numbering (dof table aliases the cells): vertex dofs are the node coordinates themselves
    // (lagrangian_basis.h:166).  A space created with explicit cells may carry a renumbered dof table (multi-GPU local
    // problems): then every dof, vertex dofs included, is located through its first incident cell.
    const bool renumbered = s->verts_p != s->dofs.p;
    if (!renumbered) {
        for (int r = 0; r < s->N; ++r)
            FDB_CUDA(cudaMemcpyAsync(out + (size_t)r * s->n_dofs, s->coords.p + (size_t)r * s->n_nodes,
                                     sizeof(double) * s->n_nodes, cudaMemcpyDeviceToDevice, s->stream));
        if (s->R == 1) return FDB_OK;
    }
    const int first_slot = renumbered ? 0 : s->M + 1;
    DevBuf<int32_t> first;
    FDB_TRY(first.alloc(s->n_dofs));
    FDB_CUDA(cudaMemsetAsync(first.p, 0x7F, sizeof(int32_t) * s->n_dofs, s->stream));
    int ns = s->nb - first_slot;
    k_first_cell<<<grid_for((int64_t)s->n_cells * ns, 256), 256, 0, s->stream>>>(s->n_cells, s->nb, first_slot, s->dofs.p,
                                                                                first.p);
    FDB_CUDA(cudaGetLastError());
    if (s->N != s->M) {
        FDB_TRY(surface_dof_coords(s, first_slot, first.p, out));
        FDB_CUDA(cudaStreamSynchronize(s->stream));
        return FDB_OK;
    }
    if (s->M == 2)
        k_edge_dof_coords<2><<<grid_for(s->n_cells, 128), 128, 0, s->stream>>>(
            s->n_cells, s->n_nodes, s->n_dofs, s->nb, first_slot, s->verts_p, s->dofs.p, s->coords.p, s->tab.p, first.p, out);
    else
        k_edge_dof_coords<3><<<grid_for(s->n_cells, 128), 128, 0, s->stream>>>(
            s->n_cells, s->n_nodes, s->n_dofs, s->nb, first_slot, s->verts_p, s->dofs.p, s->coords.p, s->tab.p, first.p, out);
    FDB_CUDA(cudaGetLastError());
    FDB_CUDA(cudaStreamSynchronize(s->stream));
    return FDB_OK;
}

int apply_dirichlet(fdb_matrix* A, const double* g, double* b, double* x0) {
    fdb_space* s = A->space;
    FDB_CHECK(A->assembled, FDB_ERR_STATE, "solver must be initialized first!");
    FDB_CHECK(s->has_boundary, FDB_ERR_STATE, "fdb_space_set_boundary has not been called");
    k_dirichlet<<<grid_for(s->n_dofs, 256), 256, 0, s->stream>>>(s->n_dofs, s->dof0_rule ? 1 : 0, A->pat->rowptr.p, A->pat->colidx.p,
                                                                s->boundary.p, g, A->val.p, b, x0);
    FDB_CUDA(cudaGetLastError());
    ++A->val_version;
    return FDB_OK;
}

}  // namespace fdb
