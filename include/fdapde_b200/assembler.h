// fdapde_b200/assembler.h -- C++17 host-side mirror of the reference's interface for the FE hot path, on top of the
// C ABI of libfdapde_b200.so (include/fdapde_b200.h).  Header only, no Eigen required.
//
// It mirrors, name for name, what the reference exposes for this path (paths relative to the fdaPDE-core tree):
//   Assembler<FEM, D, B, I>::discretize_operator / discretize_forcing   fdaPDE/finite_elements/fem_assembler.h:36-136
//   laplacian<FEM>(), diffusion<FEM>(K), advection<FEM>(b), reaction<FEM>(c), dt<FEM>()
//                                                                         fdaPDE/pde/differential_operators.h:27-38
//   operator+, operator-, unary -, double * expr, is_symmetric            fdaPDE/pde/differential_expressions.h:54-135
//   FEMSolverBase::init / set_dirichlet_bc, FEMLinearEllipticSolver::solve
//                                             fdaPDE/finite_elements/solvers/fem_solver_base.h:106-155,
//                                             fdaPDE/finite_elements/solvers/fem_linear_elliptic_solver.h:34-50
// Errors: the reference throws std::runtime_error through fdapde_assert (utils/assert.h:23-27); so does this shim
// (no exception ever crosses the C ABI itself).  When Eigen is available (FDAPDE_B200_WITH_EIGEN or
// __has_include(<Eigen/Sparse>)) SpMatrix converts to Eigen::SparseMatrix<double> with one Map + copy, which is what
// a maintainer would splice into fem_assembler.h.
#ifndef FDAPDE_B200_ASSEMBLER_H
#define FDAPDE_B200_ASSEMBLER_H

#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../fdapde_b200.h"

#if defined(FDAPDE_B200_WITH_EIGEN) || (defined(__has_include) && __has_include(<Eigen/Sparse>))
#include <Eigen/Sparse>
#define FDAPDE_B200_HAVE_EIGEN 1
#endif

namespace fdapde_b200 {

inline void check(int rc) {
    if (rc != FDB_OK) throw std::runtime_error(std::string("fdapde_b200: ") + fdb_last_error());
}

// column-major compressed matrix with int32 indices: the storage of SpMatrix<double> (utils/symbols.h:36)
struct SpMatrix {
    int rows = 0, cols = 0;
    std::vector<int32_t> outer, inner;
    std::vector<double> values;
    int64_t nonZeros() const { return (int64_t)values.size(); }
    double coeff(int i, int j) const {
        for (int32_t k = outer[j]; k < outer[j + 1]; ++k)
            if (inner[k] == i) return values[k];
        return 0.0;
    }
#ifdef FDAPDE_B200_HAVE_EIGEN
    Eigen::SparseMatrix<double> eigen() const {
        return Eigen::Map<const Eigen::SparseMatrix<double>>(rows, cols, nonZeros(), outer.data(), inner.data(),
                                                             values.data());
    }
#endif
};

// ---- operator expressions --------------------------------------------------------------------------------------------
// A flattened expression: the leaves of the reference's expression tree with the sign / scalar nodes folded in.
struct DifferentialExpr {
    struct Leaf {
        int kind;
        double scale;
        std::vector<double> coeff;
        bool space_varying;
    };
    std::vector<Leaf> leaves;
    bool is_symmetric() const {  // AND over leaves; Advection is not symmetric (advection.h:43)
        for (const Leaf& l : leaves)
            if (l.kind == FDB_ADVECTION) return false;
        return true;
    }
    void lower(fdb_opdesc* d, int symmetric = -1) const {
        if (leaves.size() > FDB_MAX_TERMS) throw std::runtime_error("fdapde_b200: too many operator terms");
        d->n_terms = (int32_t)leaves.size();
        d->symmetric = symmetric < 0 ? (is_symmetric() ? 1 : 0) : symmetric;
        for (size_t t = 0; t < leaves.size(); ++t) {
            d->terms[t].kind = leaves[t].kind;
            d->terms[t].scale = leaves[t].scale;
            d->terms[t].space_varying = leaves[t].space_varying ? 1 : 0;
            d->terms[t].coeff = leaves[t].coeff.empty() ? nullptr : leaves[t].coeff.data();
        }
    }
};
inline DifferentialExpr operator+(DifferentialExpr a, const DifferentialExpr& b) {
    a.leaves.insert(a.leaves.end(), b.leaves.begin(), b.leaves.end());
    return a;
}
inline DifferentialExpr operator-(DifferentialExpr a) {
    for (auto& l : a.leaves) l.scale = -l.scale;
    return a;
}
inline DifferentialExpr operator-(DifferentialExpr a, const DifferentialExpr& b) { return std::move(a) + (-b); }
inline DifferentialExpr operator*(double s, DifferentialExpr a) {
    for (auto& l : a.leaves) l.scale *= s;
    return a;
}
struct FEM {};  // discretisation tag (fem_symbols.h:24)
template <typename Tag = FEM> DifferentialExpr laplacian() { return {{{FDB_LAPLACIAN, 1.0, {}, false}}}; }
template <typename Tag = FEM> DifferentialExpr dt() { return {{{FDB_DT, 1.0, {}, false}}}; }
// K: N*N column-major (constant) or (n_cells*n_quad) rows of N*N column-major blocks (space varying)
template <typename Tag = FEM> DifferentialExpr diffusion(std::vector<double> K, bool space_varying = false) {
    return {{{FDB_DIFFUSION, 1.0, std::move(K), space_varying}}};
}
template <typename Tag = FEM> DifferentialExpr advection(std::vector<double> b, bool space_varying = false) {
    return {{{FDB_ADVECTION, 1.0, std::move(b), space_varying}}};
}
template <typename Tag = FEM> DifferentialExpr reaction(double c) { return {{{FDB_REACTION, 1.0, {c}, false}}}; }
template <typename Tag = FEM> DifferentialExpr reaction(std::vector<double> c) {
    return {{{FDB_REACTION, 1.0, std::move(c), true}}};
}

// ---- mesh ------------------------------------------------------------------------------------------------------------
// Triangulation<M, N> storage (triangulation.h:119-124): nodes column-major n_nodes x N, cells row-major n_cells x (M+1)
template <int M, int N> struct Triangulation {
    static constexpr int local_dim = M, embed_dim = N;
    int n_nodes = 0, n_cells = 0;
    std::vector<double> nodes;
    std::vector<int32_t> cells;
    std::vector<uint8_t> boundary;
};

// LagrangianBasis<Mesh, R>::enumerate_dofs on the device (lagrangian_basis.h:94-136)
template <int M, int N, int R> struct LagrangianBasis {
    static constexpr int n_basis = (R == 1) ? (M + 1) : (M + 1) * (M + 2) / 2;
    std::vector<int32_t> dofs;  // column-major n_cells x n_basis
    std::vector<uint8_t> boundary_dofs;
    int size = 0;
    explicit LagrangianBasis(const Triangulation<M, N>& mesh) {
        dofs.resize((size_t)mesh.n_cells * n_basis);
        boundary_dofs.resize((size_t)mesh.n_nodes + (size_t)mesh.n_cells * (M == 2 ? 3 : 6));
        check(fdb_enumerate_dofs(M, R, mesh.n_nodes, mesh.n_cells, mesh.cells.data(),
                                 mesh.boundary.empty() ? nullptr : mesh.boundary.data(), dofs.data(),
                                 boundary_dofs.data(), &size));
        boundary_dofs.resize(size);
    }
};

// Eigen's setFromTriplets on (row, col, value) lists, column-major result: duplicates are summed in list order,
// explicit zeros are kept, inner indices come out sorted (fem_assembler.h:112-113, lagrangian_basis.h:231-232).
inline SpMatrix sp_from_triplets(int rows, int cols, const std::vector<int32_t>& r, const std::vector<int32_t>& c,
                                 const std::vector<double>& v) {
    SpMatrix A;
    A.rows = rows;
    A.cols = cols;
    // stable counting sort by (col, row): equal keys keep the list order
    std::vector<int64_t> order(r.size());
    {
        std::vector<int64_t> cnt((size_t)rows + 1, 0), tmp(r.size());
        for (size_t k = 0; k < r.size(); ++k) cnt[(size_t)r[k] + 1]++;
        for (int i = 0; i < rows; ++i) cnt[(size_t)i + 1] += cnt[i];
        for (size_t k = 0; k < r.size(); ++k) tmp[(size_t)cnt[r[k]]++] = (int64_t)k;
        std::vector<int64_t> cc((size_t)cols + 1, 0);
        for (size_t k = 0; k < c.size(); ++k) cc[(size_t)c[k] + 1]++;
        for (int j = 0; j < cols; ++j) cc[(size_t)j + 1] += cc[j];
        for (size_t k = 0; k < tmp.size(); ++k) order[(size_t)cc[c[(size_t)tmp[k]]]++] = tmp[k];
    }
    A.outer.assign((size_t)cols + 1, 0);
    int64_t last = -1;
    for (size_t k = 0; k < order.size(); ++k) {
        const int64_t t = order[k];
        if (last >= 0 && c[(size_t)t] == c[(size_t)last] && r[(size_t)t] == r[(size_t)last]) {
            A.values.back() += v[(size_t)t];
        } else {
            A.inner.push_back(r[(size_t)t]);
            A.values.push_back(v[(size_t)t]);
            A.outer[(size_t)c[(size_t)t] + 1]++;
        }
        last = t;
    }
    for (int j = 0; j < cols; ++j) A.outer[(size_t)j + 1] += A.outer[j];
    return A;
}

struct SpaceDeleter { void operator()(fdb_space* s) const { fdb_space_destroy(s); } };
struct MatrixDeleter { void operator()(fdb_matrix* m) const { fdb_matrix_destroy(m); } };
struct VectorDeleter { void operator()(fdb_vector* v) const { fdb_vector_destroy(v); } };

// ---- Assembler<FEM, D, B, I> -----------------------------------------------------------------------------------------
// Same constructor arguments as fem_assembler.h:46-47 minus the integrator (the rule is implied by (M, R):
// integrator_tables.h:23-58); like the reference it keeps references to caller-owned mesh and dof table only while
// constructing (they are uploaded once).
template <int M, int N, int R> class Assembler {
   public:
    Assembler(const Triangulation<M, N>& mesh, int n_dofs, const std::vector<int32_t>& dofs_colmajor) : n_dofs_(n_dofs) {
        fdb_space* s = nullptr;
        check(fdb_space_create(&s, M, N, R, mesh.n_nodes, mesh.n_cells, mesh.nodes.data(), mesh.cells.data(), n_dofs,
                               dofs_colmajor.data()));
        space_.reset(s, SpaceDeleter());
        check(fdb_space_info(s, nullptr, &n_cells_, nullptr, &n_quad_));
    }
    // SpMatrix<double> discretize_operator(const E& op)   (fem_assembler.h:52)
    SpMatrix discretize_operator(const DifferentialExpr& op) {
        fdb_opdesc d;
        op.lower(&d);
        int64_t nnz = 0;
        check(fdb_pattern_nnz(space_.get(), d.symmetric, &nnz));
        SpMatrix A;
        A.rows = A.cols = n_dofs_;
        A.outer.resize((size_t)n_dofs_ + 1);
        A.inner.resize((size_t)nnz);
        A.values.resize((size_t)nnz);
        check(fdb_discretize_operator(space_.get(), &d, A.outer.data(), A.inner.data(), A.values.data()));
        return A;
    }
    // LagrangianBasis::eval<pointwise_evaluation>(locs) (lagrangian_basis.h:151-154, 203-235): Psi (n_locs x n_dofs,
    // [Psi]_ij = psi_j(p_i)) and the vector of ones.  locs column-major n_locs x N.
    std::pair<SpMatrix, std::vector<double>> eval_pointwise(const std::vector<double>& locs_colmajor) {
        const int64_t n = (int64_t)locs_colmajor.size() / N;
        if (n <= 0 || n * N != (int64_t)locs_colmajor.size()) throw std::runtime_error("fdapde_b200: locs must be n_locs x N");
        constexpr int nb = LagrangianBasis<M, N, R>::n_basis;
        std::vector<int32_t> cols((size_t)n * nb), rows;
        std::vector<double> vals((size_t)n * nb);
        check(fdb_eval_pointwise(space_.get(), n, locs_colmajor.data(), nullptr, cols.data(), vals.data()));
        std::vector<int32_t> r, c;
        std::vector<double> v;
        for (int64_t i = 0; i < n; ++i)
            for (int h = 0; h < nb; ++h)
                if (cols[(size_t)i * nb + h] >= 0) {   // points outside the domain have an empty row
                    r.push_back((int32_t)i);
                    c.push_back(cols[(size_t)i * nb + h]);
                    v.push_back(vals[(size_t)i * nb + h]);
                }
        return {sp_from_triplets((int)n, n_dofs_, r, c, v), std::vector<double>((size_t)n, 1.0)};
    }
    // LagrangianBasis::eval<areal_evaluation>(incidence) (lagrangian_basis.h:238-283): [Psi]_kj = int_{D_k} psi_j / |D_k|
    // and the subdomain measures.  incidence column-major n_subdomains x n_cells.
    std::pair<SpMatrix, std::vector<double>> eval_areal(const std::vector<double>& incidence_colmajor) {
        const int n_sub = (int)((int64_t)incidence_colmajor.size() / n_cells_);
        if (n_sub <= 0 || (int64_t)n_sub * n_cells_ != (int64_t)incidence_colmajor.size())
            throw std::runtime_error("fdapde_b200: incidence must be n_subdomains x n_cells");
        int64_t nt = 0;
        check(fdb_eval_areal(space_.get(), n_sub, incidence_colmajor.data(), 0, &nt, nullptr, nullptr, nullptr, nullptr));
        std::vector<int32_t> r((size_t)nt + 1), c((size_t)nt + 1);
        std::vector<double> v((size_t)nt + 1), D((size_t)n_sub);
        check(fdb_eval_areal(space_.get(), n_sub, incidence_colmajor.data(), nt + 1, &nt, r.data(), c.data(), v.data(), D.data()));
        r.resize((size_t)nt); c.resize((size_t)nt); v.resize((size_t)nt);
        return {sp_from_triplets(n_sub, n_dofs_, r, c, v), D};
    }
    // Triangulation::locate (triangulation.h:252-255)
    std::vector<int32_t> locate(const std::vector<double>& locs_colmajor) {
        const int64_t n = (int64_t)locs_colmajor.size() / N;
        std::vector<int32_t> ids((size_t)n);
        check(fdb_locate(space_.get(), n, locs_colmajor.data(), ids.data()));
        return ids;
    }
    // DVector<double> discretize_forcing(const F& f), matrix-of-values form (fem_assembler.h:122, integrator.h:85)
    std::vector<double> discretize_forcing(const std::vector<double>& f_at_quadrature_nodes) {
        if ((int64_t)f_at_quadrature_nodes.size() != (int64_t)n_cells_ * n_quad_)
            throw std::runtime_error("fdapde_b200: forcing needs n_cells * n_quad values");
        std::vector<double> b((size_t)n_dofs_);
        check(fdb_discretize_forcing(space_.get(), f_at_quadrature_nodes.data(), b.data()));
        return b;
    }
    // DVector<double> discretize_forcing(const F& f), callable form (ScalarExpr, integrator.h:80-83): f is evaluated at the
    // physical quadrature nodes J p_q + v0 -- exactly the rows of quadrature_nodes() -- and integrated like a matrix of values.
    // f takes the point as `const double*` (N coordinates).
    template <typename F, typename = decltype(std::declval<const F&>()(static_cast<const double*>(nullptr)))>
    std::vector<double> discretize_forcing(const F& f) {
        const std::vector<double> q = quadrature_nodes();
        const size_t rows = (size_t)n_cells_ * n_quad_;
        std::vector<double> fq(rows);
        double p[N];
        for (size_t k = 0; k < rows; ++k) {
            for (int d = 0; d < N; ++d) p[d] = q[(size_t)d * rows + k];
            fq[k] = f(p);
        }
        return discretize_forcing(fq);
    }
    // Integrator::quadrature_nodes(mesh) (integrator.h:109-121): column-major (n_cells * n_quad) x N
    std::vector<double> quadrature_nodes() {
        std::vector<double> q((size_t)n_cells_ * n_quad_ * N);
        check(fdb_quadrature_nodes(space_.get(), q.data()));
        return q;
    }
    fdb_space* space() const { return space_.get(); }
    int n_dofs() const { return n_dofs_; }
    int n_quadrature_nodes() const { return n_quad_; }
    int n_cells() const { return n_cells_; }

   private:
    std::shared_ptr<fdb_space> space_;  // copy-safe, as the type-erased PDE__ requires (pde.h:167-169)
    int n_dofs_ = 0, n_cells_ = 0, n_quad_ = 0;
};

// ---- FEM solver (FEMSolverBase + FEMLinearEllipticSolver) --------------------------------------------------------------
template <int M, int N, int R> class FEMLinearEllipticSolver {
   public:
    bool is_init = false;  // fem_solver_base.h:61-62
    bool success = false;
    fdb_solver_opts options{FDB_SOLVER_CG, 0, 0, 0, 1e-10};

    FEMLinearEllipticSolver(const Triangulation<M, N>& mesh) : mesh_(mesh), basis_(mesh) {
        asm_.reset(new Assembler<M, N, R>(mesh, basis_.size, basis_.dofs));
        check(fdb_space_set_boundary(asm_->space(), basis_.boundary_dofs.data()));
    }
    // FEMSolverBase::init (fem_solver_base.h:106-139): stiffness, load vector, mass
    void init(const DifferentialExpr& L, const std::vector<double>& f_at_quadrature_nodes) {
        fdb_opdesc d;
        L.lower(&d);
        options.kind = d.symmetric ? FDB_SOLVER_CG : FDB_SOLVER_BICGSTAB;
        stiff_ = make_matrix();
        check(fdb_assemble_operator(asm_->space(), &d, stiff_.get()));
        const int n = basis_.size;
        force_ = make_vector(n);
        std::shared_ptr<fdb_vector> fq = make_vector((int64_t)f_at_quadrature_nodes.size());
        check(fdb_vector_upload(fq.get(), f_at_quadrature_nodes.data(), (int64_t)f_at_quadrature_nodes.size()));
        check(fdb_assemble_forcing(asm_->space(), fq.get(), force_.get()));
        fdb_opdesc m;
        reaction<FEM>(1.0).lower(&m);
        mass_ = make_matrix();
        check(fdb_assemble_operator(asm_->space(), &m, mass_.get()));
        is_init = true;
    }
    // set_dirichlet_bc + solve (fem_solver_base.h:144-155, fem_linear_elliptic_solver.h:34-50, pde.h:102-105)
    void solve(const std::vector<double>* dirichlet_values = nullptr) {
        if (!is_init) throw std::runtime_error("solver must be initialized first!");
        const int n = basis_.size;
        std::shared_ptr<fdb_vector> x = make_vector(n);
        check(fdb_vector_fill(x.get(), 0.0));
        if (dirichlet_values) {
            std::shared_ptr<fdb_vector> g = make_vector(n);
            check(fdb_vector_upload(g.get(), dirichlet_values->data(), n));
            check(fdb_set_dirichlet(stiff_.get(), g.get(), force_.get(), x.get()));
        }
        int rc = fdb_solve(stiff_.get(), force_.get(), x.get(), &options, &stats);
        if (rc != FDB_OK && rc != FDB_ERR_NOT_CONVERGED) check(rc);
        success = (rc == FDB_OK);  // numeric failure: success = false, no throw (fem_linear_elliptic_solver.h:42-45)
        solution_.resize(n);
        check(fdb_vector_download(x.get(), solution_.data(), n));
    }
    const std::vector<double>& solution() const { return solution_; }
    SpMatrix stiff() const { return download(stiff_.get()); }
    SpMatrix mass() const { return download(mass_.get()); }
    std::vector<double> force() const {
        std::vector<double> b(basis_.size);
        check(fdb_vector_download(force_.get(), b.data(), basis_.size));
        return b;
    }
    int n_dofs() const { return basis_.size; }
    const LagrangianBasis<M, N, R>& basis() const { return basis_; }
    Assembler<M, N, R>& assembler() { return *asm_; }
    fdb_solve_stats stats{};

   private:
    std::shared_ptr<fdb_matrix> make_matrix() {
        fdb_matrix* m = nullptr;
        check(fdb_matrix_create(asm_->space(), &m));
        return std::shared_ptr<fdb_matrix>(m, MatrixDeleter());
    }
    static std::shared_ptr<fdb_vector> make_vector(int64_t n) {
        fdb_vector* v = nullptr;
        check(fdb_vector_create(n, &v));
        return std::shared_ptr<fdb_vector>(v, VectorDeleter());
    }
    SpMatrix download(fdb_matrix* m) const {
        int64_t nnz = 0;
        check(fdb_matrix_nnz(m, &nnz));
        SpMatrix A;
        A.rows = A.cols = basis_.size;
        A.outer.resize((size_t)basis_.size + 1);
        A.inner.resize((size_t)nnz);
        A.values.resize((size_t)nnz);
        check(fdb_matrix_download_csc(m, A.outer.data(), A.inner.data(), A.values.data()));
        return A;
    }
    const Triangulation<M, N>& mesh_;
    LagrangianBasis<M, N, R> basis_;
    std::shared_ptr<Assembler<M, N, R>> asm_;
    std::shared_ptr<fdb_matrix> stiff_, mass_;
    std::shared_ptr<fdb_vector> force_;
    std::vector<double> solution_;
};

// ---- FEMLinearParabolicSolver (solvers/fem_linear_parabolic_solver.h:27-72) -------------------------------------------
// dt(u) + L u = f with Dirichlet data g(., t): K = mass/dt + stiff, rows of boundary dofs replaced, then for every time
// step rhs = (mass/dt) u_i + force_{i+1} with the boundary values of t_{i+1}.  The whole loop runs on the device
// (fdb_solve_parabolic); the reference's factor-once SparseLU becomes a warm-started Krylov solve per step.
template <int M, int N, int R> class FEMLinearParabolicSolver {
   public:
    bool is_init = false;
    bool success = false;
    fdb_solver_opts options{FDB_SOLVER_CG, 0, 0, 0, 1e-12};

    // time_domain: t_0 .. t_{m-1}, equally spaced (deltaT = t_1 - t_0, fem_linear_parabolic_solver.h:43)
    FEMLinearParabolicSolver(const Triangulation<M, N>& mesh, const std::vector<double>& time_domain)
        : basis_(mesh), times_(time_domain) {
        if (times_.size() < 2) throw std::runtime_error("fdapde_b200: a parabolic problem needs at least two time instants");
        asm_.reset(new Assembler<M, N, R>(mesh, basis_.size, basis_.dofs));
        check(fdb_space_set_boundary(asm_->space(), basis_.boundary_dofs.data()));
    }
    // FEMSolverBase::init (fem_solver_base.h:106-139).  f: (n_cells * n_quad) x m column-major, the forcing at the
    // quadrature nodes for every time instant (:120-128)
    void init(const DifferentialExpr& L, const std::vector<double>& f_at_quadrature_nodes_by_time) {
        const int64_t rows = (int64_t)asm_->n_cells() * asm_->n_quadrature_nodes();
        if ((int64_t)f_at_quadrature_nodes_by_time.size() != rows * (int64_t)times_.size())
            throw std::runtime_error("fdapde_b200: forcing needs (n_cells * n_quad) x m values");
        fdb_opdesc d;
        L.lower(&d);
        options.kind = d.symmetric ? FDB_SOLVER_CG : FDB_SOLVER_BICGSTAB;
        stiff_ = make_matrix();
        check(fdb_assemble_operator(asm_->space(), &d, stiff_.get()));
        fdb_opdesc m;
        reaction<FEM>(1.0).lower(&m);
        mass_ = make_matrix();
        check(fdb_assemble_operator(asm_->space(), &m, mass_.get()));
        forcing_ = f_at_quadrature_nodes_by_time;
        is_init = true;
    }
    // solve (fem_linear_parabolic_solver.h:37-72).  initial_condition: n_dofs; dirichlet_values: n_dofs x m column-major
    // (pde.h:76) or nullptr
    void solve(const std::vector<double>& initial_condition, const std::vector<double>* dirichlet_values = nullptr) {
        if (!is_init) throw std::runtime_error("solver must be initialized first!");
        const int n = basis_.size, m = (int)times_.size();
        if ((int)initial_condition.size() != n) throw std::runtime_error("fdapde_b200: initial condition needs n_dofs values");
        if (dirichlet_values && (int64_t)dirichlet_values->size() != (int64_t)n * m)
            throw std::runtime_error("fdapde_b200: boundary data needs n_dofs x m values");
        solution_.assign((size_t)n * m, 0.0);
        int rc = fdb_solve_parabolic(stiff_.get(), mass_.get(), times_[1] - times_[0], m, forcing_.data(),
                                     dirichlet_values ? dirichlet_values->data() : nullptr, initial_condition.data(),
                                     solution_.data(), &options, &stats);
        if (rc != FDB_OK && rc != FDB_ERR_NOT_CONVERGED) check(rc);
        success = (rc == FDB_OK);
    }
    // n_dofs x m column-major, column j = u(., t_j)
    const std::vector<double>& solution() const { return solution_; }
    int n_dofs() const { return basis_.size; }
    const LagrangianBasis<M, N, R>& basis() const { return basis_; }
    Assembler<M, N, R>& assembler() { return *asm_; }
    SpMatrix stiff() const { return download(stiff_.get()); }
    SpMatrix mass() const { return download(mass_.get()); }
    // load vectors of all time instants stacked, (n_dofs * m) x 1 (fem_solver_base.h:120-128)
    std::vector<double> force() const {
        const int n = basis_.size, m = (int)times_.size();
        const size_t rows = (size_t)asm_->n_cells() * asm_->n_quadrature_nodes();
        std::vector<double> out((size_t)n * m);
        for (int j = 0; j < m; ++j) {
            std::vector<double> col(forcing_.begin() + (size_t)j * rows, forcing_.begin() + (size_t)(j + 1) * rows);
            const std::vector<double> b = asm_->discretize_forcing(col);
            std::copy(b.begin(), b.end(), out.begin() + (size_t)j * n);
        }
        return out;
    }
    fdb_solve_stats stats{};

   private:
    SpMatrix download(fdb_matrix* mm) const {
        int64_t nnz = 0;
        check(fdb_matrix_nnz(mm, &nnz));
        SpMatrix A;
        A.rows = A.cols = basis_.size;
        A.outer.resize((size_t)basis_.size + 1);
        A.inner.resize((size_t)nnz);
        A.values.resize((size_t)nnz);
        check(fdb_matrix_download_csc(mm, A.outer.data(), A.inner.data(), A.values.data()));
        return A;
    }
    std::shared_ptr<fdb_matrix> make_matrix() {
        fdb_matrix* mm = nullptr;
        check(fdb_matrix_create(asm_->space(), &mm));
        return std::shared_ptr<fdb_matrix>(mm, MatrixDeleter());
    }
    LagrangianBasis<M, N, R> basis_;
    std::vector<double> times_, forcing_, solution_;
    std::shared_ptr<Assembler<M, N, R>> asm_;
    std::shared_ptr<fdb_matrix> stiff_, mass_;
};

// ---- the reference's call shape ------------------------------------------------------------------------------------------
// Assembler<FEM, Triangulation<M, N>, LagrangianBasis<..., R>, Integrator<FEM, M, R>>(mesh, integrator, n_dofs, dofs)
// (fem_assembler.h:36-49).  The quadrature rule is implied by (M, R) (integrator_tables.h:23-58), so the integrator is a tag.
template <typename Tag, int M, int R> struct Integrator {};
namespace ref {
template <typename Tag, typename D, typename B, typename I> class Assembler;
template <int M, int N, int R>
class Assembler<FEM, Triangulation<M, N>, LagrangianBasis<M, N, R>, Integrator<FEM, M, R>> : public fdapde_b200::Assembler<M, N, R> {
   public:
    Assembler(const Triangulation<M, N>& mesh, const Integrator<FEM, M, R>&, int n_dofs, const std::vector<int32_t>& dofs)
        : fdapde_b200::Assembler<M, N, R>(mesh, n_dofs, dofs) {}
};
}  // namespace ref

// ---- PDE<D, E, F, FEM, fem_order<R>> (pde/pde.h:40-114) ----------------------------------------------------------------------
// Same members as the reference's PDE: constructors (domain, [time,] operator, forcing), set_forcing / set_differential_operator /
// set_dirichlet_bc / set_initial_condition, init(), solve(), solution() force() stiff() mass() n_dofs() dofs() dof_coords()
// quadrature_nodes().  An operator containing dt() selects the parabolic solver (pde_solver_selector, fem_solver_selector.h:29-33).
// Forcing: values at the quadrature nodes ((n_cells * n_quad) x m column-major) or, for stationary problems, a callable
// f(const double* x) (the reference has no space-time callable forcing either, fem_solver_base.h:129).
template <int M, int N, int R> class PDE {
   public:
    using Forcing = std::function<double(const double*)>;
    PDE(const Triangulation<M, N>& domain, const DifferentialExpr& diff_op, const std::vector<double>& forcing_data)
        : domain_(domain), diff_op_(diff_op), forcing_data_(forcing_data) { check_stationary(); }
    PDE(const Triangulation<M, N>& domain, const DifferentialExpr& diff_op, Forcing forcing)
        : domain_(domain), diff_op_(diff_op), forcing_fn_(std::move(forcing)) { check_stationary(); }
    PDE(const Triangulation<M, N>& domain, const std::vector<double>& time_domain, const DifferentialExpr& diff_op,
        const std::vector<double>& forcing_data)
        : domain_(domain), time_domain_(time_domain), diff_op_(diff_op), forcing_data_(forcing_data) {
        if (!is_parabolic()) throw std::runtime_error("fdapde_b200: a space-time PDE needs dt() in its operator");
    }
    // setters (pde.h:73-77)
    void set_forcing(const std::vector<double>& forcing_data) { forcing_data_ = forcing_data; forcing_fn_ = nullptr; }
    void set_forcing(Forcing f) { forcing_fn_ = std::move(f); }
    void set_differential_operator(const DifferentialExpr& op) { diff_op_ = op; }
    void set_dirichlet_bc(const std::vector<double>& data) { boundary_data_ = data; }
    void set_initial_condition(const std::vector<double>& data) { initial_condition_ = data; }
    // getters (pde.h:79-100)
    const Triangulation<M, N>& domain() const { return domain_; }
    const std::vector<double>& time_domain() const { return time_domain_; }
    const DifferentialExpr& differential_operator() const { return diff_op_; }
    const std::vector<double>& forcing_data() const { return forcing_data_; }
    const std::vector<double>& initial_condition() const { return initial_condition_; }
    const std::vector<double>& boundary_data() const { return boundary_data_; }
    int n_dofs() const { return elliptic_ ? elliptic_->n_dofs() : parabolic_->n_dofs(); }
    const std::vector<int32_t>& dofs() const { return basis().dofs; }
    const LagrangianBasis<M, N, R>& basis() const { return elliptic_ ? elliptic_->basis() : parabolic_->basis(); }
    const std::vector<double>& solution() const { return elliptic_ ? elliptic_->solution() : parabolic_->solution(); }
    std::vector<double> force() const { need_init(); return elliptic_ ? elliptic_->force() : parabolic_->force(); }
    SpMatrix stiff() const { need_init(); return elliptic_ ? elliptic_->stiff() : parabolic_->stiff(); }
    SpMatrix mass() const { need_init(); return elliptic_ ? elliptic_->mass() : parabolic_->mass(); }
    std::vector<double> quadrature_nodes() { return assembler().quadrature_nodes(); }
    std::vector<double> dof_coords() {   // column-major n_dofs x N (lagrangian_basis.h:159-183)
        std::vector<double> out((size_t)n_dofs() * N);
        check(fdb_dofs_coords(assembler().space(), out.data()));
        return out;
    }
    std::pair<SpMatrix, std::vector<double>> eval_functional_basis(const std::vector<double>& locs_colmajor) {
        return assembler().eval_pointwise(locs_colmajor);
    }
    bool success() const { return elliptic_ ? elliptic_->success : parabolic_->success; }
    bool is_init() const { return elliptic_ ? elliptic_->is_init : (parabolic_ && parabolic_->is_init); }
    // init (pde.h:101, fem_solver_base.h:106-139)
    void init() {
        if (is_parabolic()) {
            parabolic_.reset(new FEMLinearParabolicSolver<M, N, R>(domain_, time_domain_));
            parabolic_->init(diff_op_, forcing_data_);
        } else {
            elliptic_.reset(new FEMLinearEllipticSolver<M, N, R>(domain_));
            if (forcing_fn_) {
                Assembler<M, N, R>& a = elliptic_->assembler();
                const std::vector<double> q = a.quadrature_nodes();
                const size_t rows = (size_t)a.n_cells() * a.n_quadrature_nodes();
                forcing_data_.resize(rows);
                double p[N];
                for (size_t k = 0; k < rows; ++k) {
                    for (int d = 0; d < N; ++d) p[d] = q[(size_t)d * rows + k];
                    forcing_data_[k] = forcing_fn_(p);
                }
            }
            elliptic_->init(diff_op_, forcing_data_);
        }
    }
    // solve (pde.h:102-105): Dirichlet rows only when boundary data was given
    void solve() {
        need_init();
        const std::vector<double>* g = boundary_data_.empty() ? nullptr : &boundary_data_;
        if (elliptic_) elliptic_->solve(g);
        else parabolic_->solve(initial_condition_, g);
    }
    fdb_solver_opts& solver_options() { need_init(); return elliptic_ ? elliptic_->options : parabolic_->options; }

   private:
    bool is_parabolic() const {
        for (const auto& l : diff_op_.leaves) if (l.kind == FDB_DT) return true;
        return false;
    }
    void check_stationary() const {
        if (is_parabolic()) throw std::runtime_error("fdapde_b200: dt() in the operator needs the space-time constructor");
    }
    void need_init() const {
        if (!elliptic_ && !parabolic_) throw std::runtime_error("solver must be initialized first!");
    }
    Assembler<M, N, R>& assembler() { need_init(); return elliptic_ ? elliptic_->assembler() : parabolic_->assembler(); }
    const Triangulation<M, N>& domain_;
    std::vector<double> time_domain_;
    DifferentialExpr diff_op_;
    std::vector<double> forcing_data_;
    Forcing forcing_fn_;
    std::vector<double> initial_condition_, boundary_data_;
    std::shared_ptr<FEMLinearEllipticSolver<M, N, R>> elliptic_;
    std::shared_ptr<FEMLinearParabolicSolver<M, N, R>> parabolic_;
};

}  // namespace fdapde_b200

#endif  // FDAPDE_B200_ASSEMBLER_H
