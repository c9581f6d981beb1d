// C++ exercise of the drop-in boundary: the header shim include/fdapde_b200/assembler.h over libfdapde_b200.so.
// Reads like the reference's own tests (paths relative to the fdaPDE-core tree):
//   test/src/fem_operators_test.cpp:41-100   golden 6x6 P2 stiffness of -laplacian<FEM>() on c_shaped cell 175
//   test/src/fem_pde_test.cpp:43-75           P1, u = x + y, f = 0, error (mass * err^2).sum() < 1e-7
//   test/src/fem_pde_test.cpp:78-107          P2, u = 1 - x^2 - y^2, f = 4
// Build: g++ -std=c++17 -I include tests/cpp/shim_test.cpp -L fdapde-core_b200/lib -lfdapde_b200 -o shim_test
#include <cmath>
#include <cstdio>
#include <vector>

#include <cstdlib>
#include <fstream>
#include <string>

#include "fdapde_b200/assembler.h"
#include "fdapde_b200/mesh_io.h"

using namespace fdapde_b200;

static int failures = 0;
#define EXPECT_TRUE(c)                                          \
    do {                                                        \
        if (!(c)) {                                             \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            ++failures;                                         \
        }                                                       \
    } while (0)

static bool almost_equal(double a, double b, double eps = 1e-7) {  // test/src/utils/utils.h:32-41
    return std::fabs(a - b) < eps || std::fabs(a - b) < std::fmax(std::fabs(a), std::fabs(b)) * eps;
}

static Triangulation<2, 2> unit_square(int N) {  // same ordering as test/data/mesh/unit_square_16
    Triangulation<2, 2> m;
    m.n_nodes = (N + 1) * (N + 1);
    m.n_cells = 2 * N * N;
    m.nodes.resize((size_t)m.n_nodes * 2);
    m.boundary.resize(m.n_nodes);
    for (int j = 0; j <= N; ++j)
        for (int i = 0; i <= N; ++i) {
            int id = j * (N + 1) + i;
            m.nodes[id] = (double)i / N;
            m.nodes[m.n_nodes + id] = (double)j / N;
            m.boundary[id] = (i == 0 || i == N || j == 0 || j == N);
        }
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) {
            int v = j * (N + 1) + i;
            int32_t t[6] = {v, v + 1, v + N + 2, v, v + N + 2, v + N + 1};
            m.cells.insert(m.cells.end(), t, t + 6);
        }
    return m;
}

static void laplacian_order_2() {
    Triangulation<2, 2> m;
    m.n_nodes = 3;
    m.n_cells = 1;
    const double x[3] = {1.75, 1.916666666666665, 1.7918368195713161};
    const double y[3] = {0.23341855546305534, 0.267983483821244, 0.4507907046353541};
    m.nodes = {x[0], x[1], x[2], y[0], y[1], y[2]};
    m.cells = {0, 1, 2};
    std::vector<int32_t> dofs = {0, 1, 2, 3, 4, 5};
    Assembler<2, 2, 2> assembler(m, 6, dofs);
    auto L = -laplacian<FEM>();
    fdb_opdesc d;
    L.lower(&d, /*symmetric=*/0);  // the test computes all 36 integrals
    SpMatrix A;
    A.rows = A.cols = 6;
    int64_t nnz = 0;
    check(fdb_pattern_nnz(assembler.space(), 0, &nnz));
    A.outer.resize(7); A.inner.resize(nnz); A.values.resize(nnz);
    check(fdb_discretize_operator(assembler.space(), &d, A.outer.data(), A.inner.data(), A.values.data()));
    const double expected[36] = {
      0.7043890316492852,  0.1653830261033185,  0.0694133177797771, -0.6615321044132733, -0.2776532711191089,  0.0000000000000013,
      0.1653830261033185,  0.7043890316492852,  0.0694133177797769, -0.6615321044132735,  0.0000000000000003, -0.2776532711191076,
      0.0694133177797771,  0.0694133177797769,  0.4164799066786617,  0.0000000000000002, -0.2776532711191083, -0.2776532711191075,
     -0.6615321044132733, -0.6615321044132735,  0.0000000000000002,  2.4336772933029756, -0.5553065422382126, -0.5553065422382162,
     -0.2776532711191089,  0.0000000000000003, -0.2776532711191083, -0.5553065422382126,  2.4336772933029738, -1.3230642088265447,
      0.0000000000000013, -0.2776532711191075, -0.2776532711191076, -0.5553065422382162, -1.3230642088265447,  2.4336772933029751};
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) EXPECT_TRUE(almost_equal(A.coeff(i, j), expected[i * 6 + j]));
}

template <int R, typename Exact> static void poisson(int N, Exact u_ex, double f_value) {
    auto mesh = unit_square(N);
    FEMLinearEllipticSolver<2, 2, R> solver(mesh);
    solver.options.rtol = 1e-12;
    auto L = -laplacian<FEM>();
    std::vector<double> f((size_t)mesh.n_cells * solver.assembler().n_quadrature_nodes(), f_value);
    solver.init(L, f);
    // dof coordinates for the boundary data and the exact solution
    const int n = solver.n_dofs();
    std::vector<double> xy((size_t)n * 2);
    check(fdb_dofs_coords(solver.assembler().space(), xy.data()));
    std::vector<double> g(n), ex(n);
    for (int i = 0; i < n; ++i) g[i] = ex[i] = u_ex(xy[i], xy[n + i]);
    solver.solve(&g);
    EXPECT_TRUE(solver.success);
    SpMatrix Mass = solver.mass();
    double err = 0;  // (mass * err^2).sum()
    for (int j = 0; j < n; ++j)
        for (int32_t k = Mass.outer[j]; k < Mass.outer[j + 1]; ++k) {
            double e = ex[j] - solver.solution()[j];
            err += Mass.values[k] * e * e;
        }
    std::printf("P%d unit_square_%d: n_dofs %d, iterations %d, L2 error %.3e\n", R, N, n, solver.stats.iters, err);
    EXPECT_TRUE(err < 1e-7);
}

// lagrangian_basis_test.cpp:200-238 in spirit: Psi evaluated at the dof coordinates of a P2 space is the identity,
// rows of Psi sum to one at arbitrary points, areal rows average a partition of unity and D holds the areas
static void basis_evaluation() {
    auto mesh = unit_square(8);
    LagrangianBasis<2, 2, 2> basis(mesh);
    Assembler<2, 2, 2> assembler(mesh, basis.size, basis.dofs);
    const int n = basis.size;
    std::vector<double> xy((size_t)n * 2);
    check(fdb_dofs_coords(assembler.space(), xy.data()));
    auto res = assembler.eval_pointwise(xy);
    EXPECT_TRUE(res.first.rows == n && res.first.cols == n && (int)res.second.size() == n);
    for (int i = 0; i < n; ++i) EXPECT_TRUE(almost_equal(res.first.coeff(i, i), 1.0) && res.second[i] == 1.0);
    std::vector<double> pts = {0.123, 0.77, 0.5, 2.0,   // x of 4 points (the last one is outside)
                               0.456, 0.01, 0.5, 0.5};  // y
    auto psi = assembler.eval_pointwise(pts);
    for (int i = 0; i < 4; ++i) {
        double sum = 0;
        for (int j = 0; j < n; ++j) sum += psi.first.coeff(i, j);
        EXPECT_TRUE(almost_equal(sum, i < 3 ? 1.0 : 0.0, 1e-12));
    }
    auto ids = assembler.locate(pts);
    EXPECT_TRUE(ids[0] >= 0 && ids[1] >= 0 && ids[2] >= 0 && ids[3] == -1);
    // two subdomains: left half (x < 1/2) and everything
    std::vector<double> inc((size_t)2 * mesh.n_cells, 0.0);
    for (int e = 0; e < mesh.n_cells; ++e) {
        double cx = 0;
        for (int k = 0; k < 3; ++k) cx += mesh.nodes[mesh.cells[(size_t)e * 3 + k]] / 3;
        inc[(size_t)e * 2 + 0] = cx < 0.5 ? 1.0 : 0.0;
        inc[(size_t)e * 2 + 1] = 1.0;
    }
    auto ar = assembler.eval_areal(inc);
    EXPECT_TRUE(almost_equal(ar.second[0], 0.5, 1e-12) && almost_equal(ar.second[1], 1.0, 1e-12));
    for (int k = 0; k < 2; ++k) {
        double sum = 0;
        for (int j = 0; j < n; ++j) sum += ar.first.coeff(k, j);
        EXPECT_TRUE(almost_equal(sum, 1.0, 1e-12));
    }
    std::printf("basis evaluation: Psi %d x %d, %lld entries; areal %lld entries\n", res.first.rows, res.first.cols,
                (long long)res.first.nonZeros(), (long long)ar.first.nonZeros());
}

// MeshLoader round trip (test/src/utils/mesh_loader.h:62-84): files in the reference's layout -> Triangulation
static void mesh_loader() {
    auto ref = unit_square(4);
    const char* tmp = std::getenv("TMPDIR");
    const std::string dir = std::string(tmp ? tmp : "/tmp") + "/fdb_shim_mesh";
    if (std::system(("mkdir -p " + dir).c_str()) != 0) throw std::runtime_error("cannot create " + dir);
    {
        std::ofstream p(dir + "/points.csv"), e(dir + "/elements.csv"), b(dir + "/boundary.csv");
        p.precision(17);
        p << "\"\",\"V1\",\"V2\"\n";
        e << "\"\",\"V1\",\"V2\",\"V3\"\n";
        b << "\"\",\"V1\"\n";
        for (int i = 0; i < ref.n_nodes; ++i) {
            p << "\"" << i + 1 << "\",\"" << ref.nodes[i] << "\",\" " << ref.nodes[ref.n_nodes + i] << "\"\n";
            b << "\"" << i + 1 << "\"," << (int)ref.boundary[i] << "\n";
        }
        for (int c = 0; c < ref.n_cells; ++c)
            e << "\"" << c + 1 << "\"," << ref.cells[3 * c] + 1 << "," << ref.cells[3 * c + 1] + 1 << "," << ref.cells[3 * c + 2] + 1 << "\n";
    }
    auto m = load_mesh<2, 2>(dir);
    EXPECT_TRUE(m.n_nodes == ref.n_nodes && m.n_cells == ref.n_cells);
    EXPECT_TRUE(m.nodes == ref.nodes && m.cells == ref.cells && m.boundary == ref.boundary);
}

// fem_pde_test.cpp:222-285 through the C++ shim: dt(u) - lap u = f on the unit square, P2, u = sin sin exp(-t).
static void parabolic_order_2() {
    const double pi = 3.14159265358979323846;
    const int m = 11;
    std::vector<double> times(m);
    for (int j = 0; j < m; ++j) times[j] = 0.1 * j / (m - 1);
    auto mesh = unit_square(16);
    FEMLinearParabolicSolver<2, 2, 2> solver(mesh, times);
    const int n = solver.n_dofs();
    auto u = [pi](double x, double y, double t) { return std::sin(2 * pi * x) * std::sin(2 * pi * y) * std::exp(-t); };
    std::vector<double> xy((size_t)n * 2);
    check(fdb_dofs_coords(solver.assembler().space(), xy.data()));
    const int64_t nq = (int64_t)mesh.n_cells * solver.assembler().n_quadrature_nodes();
    std::vector<double> q((size_t)nq * 2), f((size_t)nq * m), g((size_t)n * m), u0(n);
    check(fdb_quadrature_nodes(solver.assembler().space(), q.data()));
    for (int j = 0; j < m; ++j) {
        for (int64_t k = 0; k < nq; ++k) f[(size_t)j * nq + k] = (8 * pi * pi - 1.0) * u(q[k], q[nq + k], times[j]);
        for (int i = 0; i < n; ++i) g[(size_t)j * n + i] = u(xy[i], xy[n + i], times[j]);
    }
    for (int i = 0; i < n; ++i) u0[i] = u(xy[i], xy[n + i], times[0]);
    solver.init(dt<FEM>() - laplacian<FEM>(), f);
    solver.solve(u0, &g);
    EXPECT_TRUE(solver.success);
    double worst = 0;
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < n; ++i) worst = std::fmax(worst, std::fabs(solver.solution()[(size_t)j * n + i] - g[(size_t)j * n + i]));
    std::printf("parabolic P2 unit_square_16: %d steps, max nodal error %.3e\n", m - 1, worst);
    EXPECT_TRUE(worst < 5e-2);  // h = 1/16: spatial error ~ (2 pi h)^3, measured 1.2e-2
}

// pde.h:40-114 through the facade: the reference's call sequence PDE(domain, L, f); set_dirichlet_bc; init; solve, with the
// forcing given as a callable (integrator.h:80-83) and as a matrix of values at the quadrature nodes (:85) -- same load vector
// bit for bit -- and the assembler spelled as in fem_assembler.h:46 (mesh, integrator, n_dofs, dofs).
static void pde_facade() {
    const double pi = 3.14159265358979323846;
    auto mesh = unit_square(32);
    auto L = -laplacian<FEM>() + reaction<FEM>(1.0);
    auto f = [pi](const double* x) { return (2 * pi * pi + 1.0) * std::sin(pi * x[0]) * std::sin(pi * x[1]); };
    PDE<2, 2, 2> pde(mesh, L, f);
    bool threw = false;
    try { pde.solve(); } catch (const std::runtime_error&) { threw = true; }   // "solver must be initialized first!"
    EXPECT_TRUE(threw);
    pde.init();
    const int n = pde.n_dofs();
    pde.set_dirichlet_bc(std::vector<double>((size_t)n, 0.0));
    pde.solver_options().rtol = 1e-12;
    pde.solve();
    EXPECT_TRUE(pde.success());
    const std::vector<double> xy = pde.dof_coords();
    double worst = 0;
    for (int i = 0; i < n; ++i) worst = std::fmax(worst, std::fabs(pde.solution()[i] - std::sin(pi * xy[i]) * std::sin(pi * xy[n + i])));
    std::printf("PDE facade P2 unit_square_32 (callable forcing): n_dofs %d, max nodal error %.3e\n", n, worst);
    EXPECT_TRUE(worst < 1e-4);
    // matrix-of-values forcing gives the same load vector; the reference-shaped assembler the same matrices
    const std::vector<double> q = pde.quadrature_nodes();
    const size_t rows = q.size() / 2;
    std::vector<double> fq(rows);
    for (size_t k = 0; k < rows; ++k) { const double x[2] = {q[k], q[rows + k]}; fq[k] = f(x); }
    LagrangianBasis<2, 2, 2> basis(mesh);
    Integrator<FEM, 2, 2> integrator;
    ref::Assembler<FEM, Triangulation<2, 2>, LagrangianBasis<2, 2, 2>, Integrator<FEM, 2, 2>> assembler(mesh, integrator, basis.size,
                                                                                                    basis.dofs);
    const std::vector<double> b_values = assembler.discretize_forcing(fq), b_callable = assembler.discretize_forcing(f);
    EXPECT_TRUE(b_values == b_callable);
    PDE<2, 2, 2> pde2(mesh, L, fq);
    pde2.init();
    EXPECT_TRUE(pde2.force() == b_values);   // (pde.force() carries the Dirichlet values after solve(), fem_solver_base.h:150)
    const SpMatrix K = assembler.discretize_operator(L), K2 = pde.stiff();
    EXPECT_TRUE(K.outer == K2.outer && K.inner == K2.inner);
    EXPECT_TRUE(K.nonZeros() == K2.nonZeros());
    // space-time: an operator with dt() selects the parabolic solver
    std::vector<double> times = {0.0, 0.01, 0.02};
    const size_t nq = rows;
    std::vector<double> ft(nq * times.size(), 0.0), u0((size_t)n, 0.0);
    PDE<2, 2, 2> heat(mesh, times, dt<FEM>() - laplacian<FEM>(), ft);
    heat.set_initial_condition(u0);
    heat.set_dirichlet_bc(std::vector<double>((size_t)n * times.size(), 0.0));
    heat.init();
    heat.solve();
    EXPECT_TRUE(heat.success());
    EXPECT_TRUE(heat.solution().size() == (size_t)n * times.size());
    EXPECT_TRUE(heat.force().size() == (size_t)n * times.size());
    double mx = 0;
    for (double v : heat.solution()) mx = std::fmax(mx, std::fabs(v));
    EXPECT_TRUE(mx == 0.0);   // zero data stay zero
    bool threw2 = false;
    try { PDE<2, 2, 2> bad(mesh, dt<FEM>() - laplacian<FEM>(), fq); } catch (const std::runtime_error&) { threw2 = true; }
    EXPECT_TRUE(threw2);
}

int main() {
    try {
        parabolic_order_2();
        pde_facade();
        laplacian_order_2();
        mesh_loader();
        basis_evaluation();
        poisson<1>(32, [](double x, double y) { return x + y; }, 0.0);
        poisson<2>(32, [](double x, double y) { return 1.0 - x * x - y * y; }, 4.0);
        // error behaviour: solve before init throws like fem_linear_elliptic_solver.h:36
        auto mesh = unit_square(4);
        FEMLinearEllipticSolver<2, 2, 1> s(mesh);
        bool threw = false;
        try { s.solve(); } catch (const std::runtime_error&) { threw = true; }
        EXPECT_TRUE(threw);
    } catch (const std::exception& e) {
        std::printf("FAILED with exception: %s\n", e.what());
        return 2;
    }
    std::printf(failures ? "SHIM_TEST_FAIL\n" : "SHIM_TEST_PASS\n");
    return failures ? 1 : 0;
}
