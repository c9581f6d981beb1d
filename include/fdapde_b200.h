/*
 * fdapde_b200.h -- C ABI of libfdapde_b200.so, the B200 (sm_100a) implementation of fdaPDE-core's
 * finite-element hot path: bilinear/linear form assembly over a Triangulation<M,N> with a Lagrange P1/P2
 * space, followed by the sparse linear solve.
 *
 * Every entry point names the reference interface it replaces (file:line relative to the fdaPDE-core tree).
 * The reference-side binding a maintainer would add is shown in INTEGRATION.md and implemented, Eigen-free,
 * in include/fdapde_b200/assembler.h.
 *
 * Conventions (identical to the reference's Eigen containers):
 *   nodes  : column-major n_nodes x N doubles            (DMatrix<double>,           triangulation.h:119)
 *   cells  : row-major    n_cells x (M+1) int32          (DMatrix<int, RowMajor>,    triangulation.h:120)
 *   dofs   : column-major n_cells x n_basis int32        (DMatrix<int>,              lagrangian_basis.h:34)
 *   sparse : column-major compressed, int32 indices, sorted inner indices, explicit zeros kept
 *            (SpMatrix<double> = Eigen::SparseMatrix<double>,                        utils/symbols.h:36)
 * All arithmetic is IEEE fp64.  No exception crosses this boundary: every call returns a status and
 * fdb_last_error() describes the last failure of the calling thread.  Handles own device memory; host arrays
 * stay caller-owned.  Calls on one handle are not thread-safe (the reference has no threading at all).
 * There is NO CPU fallback: without a CUDA device every compute call fails with FDB_ERR_CUDA.
 */
#ifndef FDAPDE_B200_H
#define FDAPDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdb_space fdb_space;   /* Triangulation<M,N> + LagrangianBasis<.,R> resident in HBM */
typedef struct fdb_matrix fdb_matrix; /* one assembled operator (values; shares the space's sparsity pattern) */
typedef struct fdb_vector fdb_vector; /* dense fp64 device vector */

enum fdb_status {
    FDB_OK = 0,
    FDB_ERR_ARG = 1,           /* fdapde_assert-like precondition failure (utils/assert.h:23-27) */
    FDB_ERR_CUDA = 2,          /* CUDA runtime failure / no device */
    FDB_ERR_STATE = 3,         /* e.g. "solver must be initialized first!" (fem_linear_elliptic_solver.h:36) */
    FDB_ERR_NOT_CONVERGED = 4, /* the reference's `success = false` (fem_linear_elliptic_solver.h:42-45) */
    FDB_ERR_UNSUPPORTED = 5
};

/* Leaves of the reference's operator expression tree (pde/differential_operators.h:33-38). */
enum fdb_term_kind { FDB_LAPLACIAN = 0, FDB_DIFFUSION = 1, FDB_ADVECTION = 2, FDB_REACTION = 3, FDB_DT = 4 };

/* One leaf with the unary minus / double* nodes above it folded into `scale`
 * (pde/differential_expressions.h:54-135).  Weak forms: laplacian.h:37-44, diffusion.h:48-55,
 * advection.h:49-56, reaction.h:47-53, dt.h:34-36.
 *   coeff, constant case     : K = N*N column-major (Eigen SMatrix), b = N, c = 1 doubles (host memory)
 *   coeff, space-varying case: one row per global quadrature node nq*e+q (integrator.h:100), row layout of
 *                              Discretized{Matrix,Vector,Scalar}Field (HOST memory, n_cells*nq rows)        */
typedef struct {
    int32_t kind;
    int32_t space_varying;
    double scale;
    const double* coeff;
} fdb_term;

#define FDB_MAX_TERMS 8
typedef struct {
    int32_t n_terms;
    int32_t symmetric; /* is_symmetric<E>: AND over leaves, Advection is false (differential_expressions.h:70-73) */
    fdb_term terms[FDB_MAX_TERMS];
} fdb_opdesc;

enum fdb_solver_kind { FDB_SOLVER_CG = 0, FDB_SOLVER_BICGSTAB = 1 };
typedef struct {
    int32_t kind;        /* fdb_solver_kind */
    int32_t jacobi;      /* 1 = diagonal preconditioner */
    int32_t maxit;       /* <= 0: 10 * n */
    int32_t check_every; /* residual is read back every this many iterations (<= 0: 32) */
    double rtol;         /* stop when ||b - A x||_2 <= rtol * ||b||_2 (Eigen's iterative-solver convention) */
} fdb_solver_opts;

typedef struct {
    int32_t iters;
    int32_t converged;
    double rel_resid; /* recurrence residual norm / ||b|| at exit */
    double seconds;   /* device time of the iteration loop (CUDA events) */
} fdb_solve_stats;

/* ---- library ---------------------------------------------------------------------------------------------- */
const char* fdb_last_error(void);
int fdb_version(void);
/* device memory released by handles is cached for reuse; fdb_trim() returns the cached blocks to the driver */
int fdb_trim(void);
/* Page-locked host memory for the arrays that cross the boundary (mesh in, CSC arrays / vectors out): the copies then
 * run at PCIe speed instead of being staged through the driver's bounce buffers.  Blocks are cached on free (pinning
 * pages is slow); fdb_trim() returns them to the OS.  Ordinary (pageable) arrays are accepted everywhere too. */
int fdb_host_alloc(size_t bytes, void** out);
int fdb_host_free(void* p);
int fdb_device_count(int* count);
int fdb_set_device(int device);

/* ---- A1/A3: mesh + DOF upload  (TriangulationBase ctor, triangulation.h:48-58; LagrangianBasis, -----------
 *      lagrangian_basis.h:94-136; Assembler ctor, fem_assembler.h:46-49)
 * M = N in {2,3}, R in {1,2}.  `cells` may be NULL when the first M+1 dof columns are the vertex ids
 * (true for every table LagrangianBasis::enumerate_dofs builds).  Uploads struct-of-arrays copies. */
int fdb_space_create(fdb_space** out, int M, int N, int R, int n_nodes, int n_cells, const double* nodes_colmajor,
                     const int32_t* cells_rowmajor, int n_dofs, const int32_t* dofs_colmajor);
void fdb_space_destroy(fdb_space* s);
/* run this space's kernels on a caller-owned cudaStream_t (default: a stream owned by the space) */
int fdb_space_set_stream(fdb_space* s, void* cuda_stream);
int fdb_space_sync(fdb_space* s);
int fdb_space_info(const fdb_space* s, int* n_dofs, int* n_cells, int* n_basis, int* n_quad);
/* per-kernel timing of fdb_assemble_operator with CUDA events on the space's stream: ms[0] = local assembly kernel,
 * ms[1] = segmented reduction kernel of the most recent assembly */
int fdb_space_set_profiling(fdb_space* s, int enabled);
int fdb_space_last_timings(fdb_space* s, double* ms, int capacity, int* count);
/* which path the last fdb_assemble_operator on this space took: *fused = 2 (k_fused_persist: persistent CTAs, every
 * block list prefetched by the bulk-copy engine, one launch), 1 (k_fused_assemble, one launch) or 0 (contribution list +
 * segmented reduction, two launches); *launches = kernels launched */
int fdb_space_last_path(const fdb_space* s, int* fused, int* launches);
/* 1 (default): fused assembly (local matrices in shared memory, no contribution list in HBM) whenever the pattern
 * admits a plan; 0: always the two-kernel path (local kernel -> sorted contribution list -> segmented reduction).
 * Both sum every entry in the same order and give bit-identical matrices. */
int fdb_space_set_fused(fdb_space* s, int enabled);
/* boundary dof markers, BinaryVector<Dynamic> boundary_dofs_ (fem_solver_base.h:102), one byte per dof */
int fdb_space_set_boundary(fdb_space* s, const uint8_t* boundary_dofs);
/* The reference's boundary iterator always visits dof 0 (fem_solver_base.h:86); this is mirrored
 * by default.  A rank of a partitioned problem whose local dof 0 is not global dof 0 must switch the rule off. */
int fdb_space_set_dof0_rule(fdb_space* s, int enabled);

/* A2+A3 on the device: LagrangianBasis::enumerate_dofs (lagrangian_basis.h:94-136) on top of the edge numbering
 * of Triangulation<2,N> / Triangulation<3,3> (triangulation.h:143-196, 319-399), by sort/unique instead of hash
 * maps.  dofs_colmajor: n_cells x n_basis; boundary_dofs: capacity n_nodes + n_cells*(M==2?3:6) bytes. */
int fdb_enumerate_dofs(int M, int R, int n_nodes, int n_cells, const int32_t* cells_rowmajor,
                       const uint8_t* boundary_nodes, int32_t* dofs_colmajor, uint8_t* boundary_dofs, int* n_dofs);

/* ---- N3: the topology the Triangulation constructors build (geometry/triangulation.h:143-196 triangles, incl. the
 * surface case Triangulation<2,3>; :319-399 tetrahedra), on the device by sort/unique instead of std::unordered_map scans.
 * Ids reproduce the reference's first-occurrence numbering exactly.  "Facet" = edge of a triangle / face of a tetrahedron.
 *   neighbors       n_cells x (M+1) row-major: cell across the facet opposite to local vertex j, -1 on the boundary
 *                   (TriangulationBase::neighbors(), triangulation.h:57,180-181,360-361)
 *   facets          n_facets x M sorted node ids            (2D: edges(), 3D: faces())
 *   cell_to_facets  n_cells x (M+1)                          (2D: cell_to_edges(), 3D: cell_to_faces())
 *   facet_to_cells  n_facets x 2, second = -1 on the boundary (2D: edge_to_cells(), 3D: face_to_cells())
 *   facet_boundary  n_facets                                 (2D: boundary_edges(), 3D: boundary_faces())
 *   3D only: edges n_edges x 2 (numbered inside every new face, :356-373), face_to_edges n_faces x 3,
 *            edge_boundary (both end nodes on the boundary, :370), edge_to_cells as ascending lists
 *            edge_cells[edge_cell_ptr[e] .. edge_cell_ptr[e+1])  (an unordered_set per edge in the reference, :486)
 * Any output pointer of fdb_topology_download may be NULL. */
typedef struct fdb_topology fdb_topology;
int fdb_topology_create(fdb_topology** out, int M, int n_nodes, int n_cells, const int32_t* cells_rowmajor,
                        const uint8_t* boundary_nodes);
void fdb_topology_destroy(fdb_topology* t);
int fdb_topology_sizes(const fdb_topology* t, int* n_facets, int* n_edges, int64_t* n_edge_cells);
int fdb_topology_download(const fdb_topology* t, int32_t* neighbors, int32_t* facets, int32_t* cell_to_facets,
                          int32_t* facet_to_cells, uint8_t* facet_boundary, int32_t* edges, int32_t* face_to_edges,
                          uint8_t* edge_boundary, int32_t* edge_cell_ptr, int32_t* edge_cells);

/* Integrator::quadrature_nodes (integrator.h:109-121): column-major (n_cells*nq) x N, row nq*e+q */
int fdb_quadrature_nodes(fdb_space* s, double* out_colmajor);
/* LagrangianBasis::dofs_coords (lagrangian_basis.h:159-183): column-major n_dofs x N */
int fdb_dofs_coords(fdb_space* s, double* out_colmajor);

/* ---- point location and basis evaluation matrices Psi ---------------------------------------
 * Triangulation::locate (triangulation.h:252-255, tree_search.h:71-86 + Simplex::contains, simplex.h:115-128):
 * cell_ids[i] = a cell containing point i (locs_colmajor: n_locs x N), -1 if none.  A point shared by several
 * cells gets the smallest cell id (the reference's choice follows std::unordered_set iteration order). */
int fdb_locate(fdb_space* s, int64_t n_locs, const double* locs_colmajor, int32_t* cell_ids);
/* pointwise_evaluation<LagrangianBasis>::eval (lagrangian_basis.h:203-235): the triplet list of Psi in emission
 * order, n_basis slots per point: (i, cols[i*nb + h], vals[i*nb + h]) = (i, dofs(e, h), psi_h(p_i)); cols = -1 and
 * vals = 0 for a point outside the domain (its row of Psi is empty).  cell_ids may be NULL.  The second member of the
 * reference's pair (a vector of ones) is left to the caller. */
int fdb_eval_pointwise(fdb_space* s, int64_t n_locs, const double* locs_colmajor, int32_t* cell_ids, int32_t* cols,
                       double* vals);
/* areal_evaluation<LagrangianBasis>::eval (lagrangian_basis.h:238-283): incidence_colmajor is n_subdomains x n_cells
 * (== 1: the cell belongs to the subdomain).  Output: the reference's triplet list in emission order (subdomain,
 * cells ascending, basis function), values already divided by the subdomain measure, duplicates left to
 * setFromTriplets; measures[k] = D_k.  n_triplets = n_basis * number of ones; with all four output arrays NULL only
 * n_triplets is computed (size query), otherwise capacity >= n_triplets is required. */
int fdb_eval_areal(fdb_space* s, int n_subdomains, const double* incidence_colmajor, int64_t capacity,
                   int64_t* n_triplets, int32_t* rows, int32_t* cols, double* vals, double* measures);

/* ---- sparsity pattern + scatter map (built once per space and symmetry class, on the device) --------------
 * what setFromTriplets/makeCompressed/selfadjointView produce structurally (fem_assembler.h:112-117) */
int fdb_pattern_nnz(fdb_space* s, int symmetric, int64_t* nnz);
/* Builds the pattern AND the fused-assembly plan now.  Without this call the plan is built lazily by the second
 * assembly on a pattern: a one-shot discretize_operator is cheaper on the two-kernel path. */
int fdb_space_prepare(fdb_space* s, int symmetric);
int fdb_pattern_download(fdb_space* s, int symmetric, int32_t* outer, int32_t* inner);

/* ---- A8: Assembler::discretize_operator (fem_assembler.h:52-121) ------------------------------------------ */
int fdb_matrix_create(fdb_space* s, fdb_matrix** out);
void fdb_matrix_destroy(fdb_matrix* A);
int fdb_matrix_nnz(const fdb_matrix* A, int64_t* nnz);
/* device-resident assembly: local matrices in registers + deterministic segmented reduction, no atomics */
int fdb_assemble_operator(fdb_space* s, const fdb_opdesc* op, fdb_matrix* A);
/* CSC arrays exactly as the reference's SpMatrix<double> holds them (outer n_dofs+1, inner nnz, values nnz) */
int fdb_matrix_download_csc(fdb_matrix* A, int32_t* outer, int32_t* inner, double* values);
/* host-buffer convenience = fdb_assemble_operator + fdb_matrix_download_csc (the call the header shim makes) */
int fdb_discretize_operator(fdb_space* s, const fdb_opdesc* op, int32_t* outer, int32_t* inner, double* values);

/* ---- dense vectors ----------------------------------------------------------------------------------------- */
int fdb_vector_create(int64_t n, fdb_vector** out);
void fdb_vector_destroy(fdb_vector* v);
int fdb_vector_upload(fdb_vector* v, const double* host, int64_t n);
int fdb_vector_download(const fdb_vector* v, double* host, int64_t n);
int fdb_vector_fill(fdb_vector* v, double value);

/* ---- A9: Assembler::discretize_forcing (fem_assembler.h:122-136, integrator.h:74-90) ---------------------- */
/* f_quad: n_cells*nq values of the forcing at the quadrature nodes (row nq*e+q); b: n_dofs */
int fdb_assemble_forcing(fdb_space* s, const fdb_vector* f_quad, fdb_vector* b);
int fdb_discretize_forcing(fdb_space* s, const double* f_quad_host, double* b_host);

/* ---- A9c: FEMSolverBase::set_dirichlet_bc (fem_solver_base.h:144-155) ------------------------------------- */
/* rows of boundary dofs (and always dof 0, :86) <- unit rows, b(d) <- g(d); x0 (optional) gets g(d) on those
 * rows so that CG on the row-replaced matrix is CG on the SPD interior block.  Needs fdb_space_set_boundary. */
int fdb_set_dirichlet(fdb_matrix* A, const fdb_vector* g, fdb_vector* b, fdb_vector* x0);

/* ---- A9d: FEMLinearEllipticSolver::solve (fem_linear_elliptic_solver.h:34-50) ------------------------------ */
/* The reference's SparseLU is replaced by fused CG (SPD) / BiCGSTAB (non-symmetric).  x: in = initial guess
 * (must hold the Dirichlet values), out = solution.  Returns FDB_ERR_NOT_CONVERGED like `success = false`. */
int fdb_solve(fdb_matrix* A, const fdb_vector* b, fdb_vector* x, const fdb_solver_opts* opts,
              fdb_solve_stats* stats);
int fdb_solve_host(fdb_matrix* A, const double* b_host, double* x_host, const fdb_solver_opts* opts,
                   fdb_solve_stats* stats);
/* ---- N2: FEMLinearParabolicSolver::solve (solvers/fem_linear_parabolic_solver.h:37-72) ------------------------------
 * K = mass/dt + stiff with Dirichlet rows replaced (:49-56), then for i = 0..m-2 (:64-70):
 *     rhs = (mass/dt) u_i + force_{i+1};  rhs(d) = g(d, i+1) on boundary dofs;  u_{i+1} = K^-1 rhs
 * (the comment in the reference says forward Euler; the algebra is implicit Euler).  The whole time loop runs on
 * the device; the reference's SparseLU factor-once/solve-many becomes CG/BiCGSTAB warm-started from u_i.
 *   f_quad   : (n_cells*nq) x m column-major, forcing at the quadrature nodes per time step (fem_solver_base.h:120-128)
 *   g        : n_dofs x m column-major boundary data (pde.h:76), may be NULL (no Dirichlet rows)
 *   u0       : n_dofs initial condition;  solution: n_dofs x m column-major, column 0 = u0
 * stiff and mass must be assembled on the same space; stiff is NOT modified.
 * A step that misses the tolerance does not end the loop: every column is filled (the reference solves every step
 * too), stats->converged = 0 and the call returns FDB_ERR_NOT_CONVERGED.  With g == NULL no row of K is replaced --
 * a divergence from the reference, which always zeroes the boundary rows (:49-56); pass a zero g for that. */
int fdb_solve_parabolic(fdb_matrix* stiff, fdb_matrix* mass, double dt, int m, const double* f_quad, const double* g,
                        const double* u0, double* solution, const fdb_solver_opts* opts, fdb_solve_stats* stats);
/* C = a*A + b*B on matrices of one space (values only; the structural pattern is shared) */
int fdb_matrix_axpby(fdb_matrix* C, double a, const fdb_matrix* A, double b, const fdb_matrix* B);

/* CG as one persistent cooperative kernel: 0 never, 1 only for partitioned matrices with a peer-memory plan (default),
 * 2 always.  Both forms run the same recurrences with the same deterministic reductions. */
int fdb_set_persistent_cg(int mode);
/* CG (single-reduction recurrence) and BiCGSTAB as one persistent sliced-ELL kernel per GPU, halo exchange and
 * reductions over NVLink peer memory (csrc/solve_peer.cu): 0 never, 1 for partitioned matrices with a peer-memory plan
 * (default), 2 also on a single GPU. */
int fdb_set_persistent_sell(int mode);
/* y = A x  (building block of the solvers, exposed for verification and the roofline measurement) */
int fdb_spmv(fdb_matrix* A, const fdb_vector* x, fdb_vector* y);

/* ---- multi-GPU solve (no counterpart in the reference, which is single process / single thread) -----------------
 * One process per GPU.  Each rank creates its space on the LOCAL mesh: the cells touching its owned dof rows, with
 * local numbering [owned dofs | halo dofs grouped by owning neighbour rank].  Assembly then needs no communication and
 * every owned row is summed in the same cell order as on one GPU.  fdb_matrix_set_partition turns fdb_solve / fdb_spmv
 * on that matrix into their distributed versions: NCCL halo exchange before each SpMV, fp64 all-reduce for the dots.
 * The 128-byte id comes from fdb_comm_unique_id on one rank and is broadcast by the host (e.g. torch.distributed). */
typedef struct fdb_comm fdb_comm;
int fdb_comm_unique_id(void* id128);
int fdb_comm_create(fdb_comm** out, int rank, int world_size, const void* id128);
void fdb_comm_destroy(fdb_comm* c);
/* neighbor_ranks[n_neighbors]; send_idx = owned local indices to send, grouped by neighbour (send_counts);
 * recv_counts = halo dofs received from each neighbour, in the order they are numbered after the owned dofs */
int fdb_matrix_set_partition(fdb_matrix* A, fdb_comm* comm, int n_owned, int n_neighbors, const int32_t* neighbor_ranks,
                             const int32_t* send_counts, const int32_t* send_idx, const int32_t* recv_counts);

/* Peer-memory plan for the persistent multi-GPU CG (one cooperative kernel per rank; halo values and reduction
 * operands are stored straight over NVLink into buffers of the other ranks mapped with cudaIpc, each value carrying
 * its own arrival tag -- no NCCL call, fence or flag inside the loop).
 *   1. every rank: fdb_matrix_peer_export -> 64-byte cudaIpcMemHandle of its exchange buffer
 *   2. host all-gathers the handles and every rank's halo size
 *   3. every rank: fdb_matrix_peer_connect with, for each of ITS neighbours (same order as fdb_matrix_set_partition),
 *      the offset inside that neighbour's halo where this rank's entries go (the neighbour's receive offset for it).
 * Without these two calls a partitioned fdb_solve uses the NCCL multi-kernel loop. */
int fdb_matrix_peer_export(fdb_matrix* A, void* ipc_handle64);
int fdb_matrix_peer_connect(fdb_matrix* A, const void* handles64, const int64_t* peer_n_halo,
                            const int32_t* nbr_recv_offset);

#ifdef __cplusplus
}
#endif
#endif /* FDAPDE_B200_H */
