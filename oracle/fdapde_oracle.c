/*
 * fdapde_oracle.c -- CPU restatement of fdaPDE-core's finite-element hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA product in
 * fdapde-core_b200/: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  Nothing under
 * fdapde-core_b200/ links, imports or executes it.
 *
 * The reference (fdaPDE/fdaPDE-core, C++20 header only) cannot be compiled in
 * this image: its arithmetic partly lives in Eigen 3.4.0 (README.md:19,
 * test/CMakeLists.txt:9), which is neither vendored under /root/reference nor
 * installed.  This file therefore restates, single-threaded and in plain C,
 *   - the reference's own code for the path (file:line cited at every function,
 *     paths relative to /root/reference), and
 *   - the published algorithms of the Eigen 3.4.0 call sites on the path
 *     (setFromTriplets, selfadjointView<Lower>, fixed-size inverse/determinant,
 *     PartialPivLU), restated from Eigen's documented behaviour.
 *
 * Parity pinning: checked in tests/test_oracle_golden.py against every golden
 * vector the reference's tests hold for this path (fem_operators_test.cpp:83-96,
 * lagrangian_basis_test.cpp:111-114,133-140,158-161,184-187, simplex_test.cpp:34,97,
 * integration_test.cpp:46-80, fem_pde_test.cpp:74,106,165,211, mesh_loader.h:35).
 * The reference holds no golden vector for a *global* matrix, pattern or solution
 * vector: for those, parity is "unpinned" by the reference's own
 * tests and rests on Eigen's documented setFromTriplets semantics restated here.
 *
 * Conventions: nodes are column-major n_nodes x N (Eigen DMatrix<double>),
 * cells row-major n_cells x (M+1) (triangulation.h:121), dofs column-major
 * n_cells x nb (lagrangian_basis.h:34), all indices int32, all arithmetic fp64.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXNB 10
#define ORC_MAXNQ 12
#define ORC_MAXD 3

enum { ORC_LAPLACIAN = 0, ORC_DIFFUSION = 1, ORC_ADVECTION = 2, ORC_REACTION = 3, ORC_DT = 4 };

/* One leaf of the operator expression tree (pde/differential_expressions.h:54-135), with the
 * product of the unary minus / double* nodes above it folded into `scale`.  coeff: constant case
 * = K (N*N, column-major like Eigen SMatrix), b (N) or c (1); space-varying case = one row per
 * global quadrature node nq*e+q (integrator.h:100), row layout as DiscretizedMatrixField
 * (fields/matrix_expressions.h:207-221: column-major N*N inside the row),
 * DiscretizedVectorField (vector_expressions.h:103-115) and DiscretizedScalarField
 * (scalar_expressions.h:98-108). */
typedef struct {
    int32_t kind;
    int32_t space_varying;
    double scale;
    const double* coeff;
} orc_term;

/* ------------------------------------------------------------------------------------------
 * combinatorics / sizes
 * ---------------------------------------------------------------------------------------- */
static int ct_factorial(int n) { return n ? n * ct_factorial(n - 1) : 1; } /* utils/combinatorics.h:29 */
static int ct_binomial(int n, int m) { return ct_factorial(n) / (ct_factorial(m) * ct_factorial(n - m)); }

int orc_n_basis(int M, int R) { return ct_binomial(M + R, R); } /* lagrangian_basis.h:44 */

/* integrator_tables.h:23-58 standard_fem_quadrature_rule */
int orc_n_quad(int M, int R) {
    switch (M) {
    case 1: return R == 1 ? 2 : 3;
    case 2: return R == 1 ? 3 : (R == 2 ? 6 : 12);
    case 3: return R == 1 ? 4 : 5;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * quadrature tables: constants verbatim from utils/integration/integrator_tables.h:64-320
 * (they are data: 15-16 significant digits as coded, e.g. 3 x 0.333333333333333)
 * ---------------------------------------------------------------------------------------- */
int orc_quadrature_table(int M, int K, double* nodes /* K*M */, double* weights /* K */) {
    static const double n1_2[] = {0.211324865405187, 0.788675134594812};
    static const double w1_2[] = {0.500000000000000, 0.500000000000000};
    static const double n1_3[] = {0.112701665379258, 0.500000000000000, 0.887298334620741};
    static const double w1_3[] = {0.277777777777778, 0.444444444444444, 0.277777777777778};
    static const double n2_1[] = {0.333333333333333, 0.333333333333333};
    static const double w2_1[] = {1.};
    static const double n2_3[] = {0.166666666666667, 0.166666666666667, 0.666666666666667,
                                  0.166666666666667, 0.166666666666667, 0.666666666666667};
    static const double w2_3[] = {0.333333333333333, 0.333333333333333, 0.333333333333333};
    static const double n2_6[] = {0.445948490915965, 0.445948490915965, 0.445948490915965, 0.108103018168070,
                                  0.108103018168070, 0.445948490915965, 0.091576213509771, 0.091576213509771,
                                  0.091576213509771, 0.816847572980459, 0.816847572980459, 0.091576213509771};
    static const double w2_6[] = {0.223381589678011, 0.223381589678011, 0.223381589678011,
                                  0.109951743655322, 0.109951743655322, 0.109951743655322};
    static const double n2_7[] = {0.333333333333333, 0.333333333333333, 0.101286507323456, 0.101286507323456,
                                  0.101286507323456, 0.797426985353087, 0.797426985353087, 0.101286507323456,
                                  0.470142064105115, 0.470142064105115, 0.470142064105115, 0.059715871789770,
                                  0.059715871789770, 0.470142064105115};
    static const double w2_7[] = {0.225000000000000, 0.125939180544827, 0.125939180544827, 0.125939180544827,
                                  0.132394152788506, 0.132394152788506, 0.132394152788506};
    static const double n2_12[] = {0.873821971016996, 0.063089014491502, 0.063089014491502, 0.873821971016996,
                                   0.063089014491502, 0.063089014491502, 0.501426509658179, 0.249286745170910,
                                   0.249286745170910, 0.501426509658179, 0.249286745170910, 0.249286745170910,
                                   0.636502499121399, 0.310352451033785, 0.636502499121399, 0.053145049844816,
                                   0.310352451033785, 0.636502499121399, 0.310352451033785, 0.053145049844816,
                                   0.053145049844816, 0.636502499121399, 0.053145049844816, 0.310352451033785};
    static const double w2_12[] = {0.050844906370207, 0.050844906370207, 0.050844906370207, 0.116786275726379,
                                   0.116786275726379, 0.116786275726379, 0.082851075618374, 0.082851075618374,
                                   0.082851075618374, 0.082851075618374, 0.082851075618374, 0.082851075618374};
    static const double n3_1[] = {0.250000000000000, 0.250000000000000, 0.250000000000000};
    static const double w3_1[] = {1.};
    static const double n3_4[] = {0.585410196624969, 0.138196601125011, 0.138196601125011, 0.138196601125011,
                                  0.138196601125011, 0.138196601125011, 0.138196601125011, 0.138196601125011,
                                  0.585410196624969, 0.138196601125011, 0.585410196624969, 0.138196601125011};
    static const double w3_4[] = {0.250000000000000, 0.250000000000000, 0.250000000000000, 0.250000000000000};
    static const double n3_5[] = {0.250000000000000, 0.250000000000000, 0.250000000000000, 0.500000000000000,
                                  0.166666666666667, 0.166666666666667, 0.166666666666667, 0.500000000000000,
                                  0.166666666666667, 0.166666666666667, 0.166666666666667, 0.500000000000000,
                                  0.166666666666667, 0.166666666666667, 0.166666666666667};
    static const double w3_5[] = {-0.80000000000000, 0.450000000000000, 0.450000000000000, 0.450000000000000,
                                  0.450000000000000};
    static const double n3_11[] = {
      0.2500000000000000, 0.2500000000000000, 0.2500000000000000, 0.7857142857142857, 0.0714285714285714,
      0.0714285714285714, 0.0714285714285714, 0.0714285714285714, 0.0714285714285714, 0.0714285714285714,
      0.0714285714285714, 0.7857142857142857, 0.0714285714285714, 0.7857142857142857, 0.0714285714285714,
      0.1005964238332008, 0.3994035761667992, 0.3994035761667992, 0.3994035761667992, 0.1005964238332008,
      0.3994035761667992, 0.3994035761667992, 0.3994035761667992, 0.1005964238332008, 0.3994035761667992,
      0.1005964238332008, 0.1005964238332008, 0.1005964238332008, 0.3994035761667992, 0.1005964238332008,
      0.1005964238332008, 0.1005964238332008, 0.3994035761667992};
    static const double w3_11[] = {-0.0789333333333333, 0.0457333333333333, 0.0457333333333333, 0.0457333333333333,
                                   0.0457333333333333,  0.1493333333333333, 0.1493333333333333, 0.1493333333333333,
                                   0.1493333333333333,  0.1493333333333333, 0.1493333333333333};
    const double *n = 0, *w = 0;
#define ORC_PICK(m, k, nn, ww) if (M == m && K == k) { n = nn; w = ww; }
    ORC_PICK(1, 2, n1_2, w1_2) ORC_PICK(1, 3, n1_3, w1_3) ORC_PICK(2, 1, n2_1, w2_1) ORC_PICK(2, 3, n2_3, w2_3)
    ORC_PICK(2, 6, n2_6, w2_6) ORC_PICK(2, 7, n2_7, w2_7) ORC_PICK(2, 12, n2_12, w2_12) ORC_PICK(3, 1, n3_1, w3_1)
    ORC_PICK(3, 4, n3_4, w3_4) ORC_PICK(3, 5, n3_5, w3_5) ORC_PICK(3, 11, n3_11, w3_11)
#undef ORC_PICK
    if (!n) return -1;
    memcpy(nodes, n, sizeof(double) * (size_t)(K * M));
    memcpy(weights, w, sizeof(double) * (size_t)K);
    return 0;
}

/* Integrator<FEM,M,R> (integrator.h:36-43): the standard rule for (M,R) */
int orc_quadrature(int M, int R, double* nodes, double* weights) {
    return orc_quadrature_table(M, orc_n_quad(M, R), nodes, weights);
}

/* ------------------------------------------------------------------------------------------
 * reference element + Lagrange basis
 * ---------------------------------------------------------------------------------------- */
/* basis/reference_element.h:28-97 ReferenceElement<M,R>::nodes */
int orc_reference_nodes(int M, int R, double* nodes /* nb*M */) {
    static const double r11[] = {0, 1};
    static const double r12[] = {0, 1, 0.5};
    static const double r21[] = {0, 0, 1, 0, 0, 1};
    static const double r22[] = {0, 0, 1, 0, 0, 1, 0.5, 0, 0, 0.5, 0.5, 0.5};
    static const double r31[] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
    static const double r32[] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0.5, 0.5, 0, 0, 0.5, 0, 0.5, 0, 0,
                                 0.5, 0, 0.5, 0, 0.5, 0.5, 0, 0, 0.5};
    const double* r = 0;
    if (M == 1 && R == 1) r = r11;
    if (M == 1 && R == 2) r = r12;
    if (M == 2 && R == 1) r = r21;
    if (M == 2 && R == 2) r = r22;
    if (M == 3 && R == 1) r = r31;
    if (M == 3 && R == 2) r = r32;
    if (!r) return -1;
    memcpy(nodes, r, sizeof(double) * (size_t)(orc_n_basis(M, R) * M));
    return 0;
}

/* basis/multivariate_polynomial.h:52-79 ct_poly_exp: exponent table, first coordinate fastest */
void orc_poly_table(int N, int R, int* table /* nmon*N */) {
    int monomials = ct_binomial(R + N, R);
    int tmp[ORC_MAXD + 1] = {0, 0, 0, 0};
    int j = 0;
    while (j < monomials) {
        int i = 0, found = 0;
        while (i < N && !found) {
            int sum = 0;
            for (int k = 0; k < N; ++k) sum += tmp[k];
            if (tmp[i] <= R && sum <= R) {
                found = 1;
                for (int k = 0; k < N; ++k) table[j * N + k] = tmp[k];
                tmp[0]++;
                j++;
            } else {
                tmp[i] = 0;
                tmp[++i]++;
            }
        }
    }
}

/* multivariate_polynomial.h:111-125 MonomialProduct: x_{N-1}^{e_{N-1}} * ( ... * x_0^{e_0}) with std::pow */
static double monomial_product(int N, const double* p, const int* e) {
    double m = e[0] == 0 ? 1 : pow(p[0], e[0]);
    for (int k = 1; k < N; ++k)
        if (e[k] != 0) m = pow(p[k], e[k]) * m;
    return m;
}

/* multivariate_polynomial.h:127-145,209-213: p(x) = sum_m c_m * monomial_m(x), summed from m = 0 upward */
double orc_poly_eval(int N, int R, const double* coeff, const double* p) {
    int nm = ct_binomial(R + N, R);
    int table[ORC_MAXNB * ORC_MAXD];
    orc_poly_table(N, R, table);
    double v = coeff[0] * monomial_product(N, p, table);
    for (int m = 1; m < nm; ++m) v = (coeff[m] * monomial_product(N, p, table + m * N)) + v;
    return v;
}

/* multivariate_polynomial.h:81-108 (ct_grad_exp) + :172-182 PolynomialDerivative::operator() */
double orc_poly_grad(int N, int R, const double* coeff, int dir, const double* p) {
    int nm = ct_binomial(R + N, R);
    int table[ORC_MAXNB * ORC_MAXD];
    orc_poly_table(N, R, table);
    double value = 0;
    for (int m = 0; m < nm; ++m) {
        if (table[m * N + dir] != 0) {
            int ge[ORC_MAXD];
            for (int z = 0; z < N; ++z)
                ge[z] = (z == dir) ? (table[m * N + z] == 0 ? 0 : table[m * N + z] - 1) : table[m * N + z];
            value += coeff[m] * table[m * N + dir] * monomial_product(N, p, ge);
        }
    }
    return value;
}

/* Eigen 3.4.0 PartialPivLU (unblocked, right-looking, row pivoting on max |.|) as used by
 * lagrangian_basis.h:78-90: solves V a = e_i for every i.  coeff[i*nb + m] = m-th monomial
 * coefficient of basis function i. */
void orc_ref_basis_coeffs(int M, int R, double* coeff /* nb*nb */) {
    int nb = orc_n_basis(M, R);
    double nodes[ORC_MAXNB * ORC_MAXD];
    int table[ORC_MAXNB * ORC_MAXD];
    double lu[ORC_MAXNB * ORC_MAXNB];
    int perm[ORC_MAXNB];
    orc_reference_nodes(M, R, nodes);
    orc_poly_table(M, R, table);
    /* lagrangian_basis.h:69-76 Vandermonde matrix, column 0 = ones */
    for (int i = 0; i < nb; ++i) {
        lu[i * nb + 0] = 1.0;
        for (int j = 1; j < nb; ++j) lu[i * nb + j] = monomial_product(M, nodes + i * M, table + j * M);
    }
    for (int i = 0; i < nb; ++i) perm[i] = i;
    for (int k = 0; k < nb; ++k) {
        int piv = k;
        double best = fabs(lu[k * nb + k]);
        for (int r = k + 1; r < nb; ++r)
            if (fabs(lu[r * nb + k]) > best) { best = fabs(lu[r * nb + k]); piv = r; }
        if (piv != k) {
            for (int c = 0; c < nb; ++c) { double t = lu[k * nb + c]; lu[k * nb + c] = lu[piv * nb + c]; lu[piv * nb + c] = t; }
            int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
        }
        for (int r = k + 1; r < nb; ++r) lu[r * nb + k] /= lu[k * nb + k];
        for (int r = k + 1; r < nb; ++r)
            for (int c = k + 1; c < nb; ++c) lu[r * nb + c] -= lu[r * nb + k] * lu[k * nb + c];
    }
    for (int i = 0; i < nb; ++i) {
        double a[ORC_MAXNB];
        for (int r = 0; r < nb; ++r) a[r] = (perm[r] == i) ? 1.0 : 0.0;
        for (int r = 0; r < nb; ++r)
            for (int c = 0; c < r; ++c) a[r] -= lu[r * nb + c] * a[c];
        for (int r = nb - 1; r >= 0; --r) {
            for (int c = r + 1; c < nb; ++c) a[r] -= lu[r * nb + c] * a[c];
            a[r] /= lu[r * nb + r];
        }
        for (int m = 0; m < nb; ++m) coeff[i * nb + m] = a[m];
    }
}

/* ------------------------------------------------------------------------------------------
 * geometry (geometry/simplex.h:184-195; Eigen 3.4.0 fixed-size inverse()/determinant())
 * v: (M+1) x N vertex coordinates, vertex-major.  J, invJ: row-major N x M and M x N.
 * ---------------------------------------------------------------------------------------- */
static double cof3(const double* m, int i, int j) { /* Eigen cofactor_3x3<i,j> */
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
static double det3_helper(const double* m, int a, int b, int c) { /* Eigen bruteforce_det3_helper */
    return m[0 * 3 + a] * (m[1 * 3 + b] * m[2 * 3 + c] - m[1 * 3 + c] * m[2 * 3 + b]);
}

int orc_cell_geometry(int M, int N, const double* v, double* J, double* invJ, double* measure) {
    for (int j = 0; j < M; ++j)
        for (int r = 0; r < N; ++r) J[r * M + j] = v[(j + 1) * N + r] - v[r]; /* simplex.h:185 */
    if (M == N && M == 2) {
        double det = J[0] * J[3] - J[2] * J[1];
        double invdet = 1.0 / det;
        invJ[0] = J[3] * invdet;
        invJ[2] = -J[2] * invdet;
        invJ[1] = -J[1] * invdet;
        invJ[3] = J[0] * invdet;
        *measure = fabs(det) / 2; /* simplex.h:188 */
        return 0;
    }
    if (M == N && M == 3) {
        double c0[3] = {cof3(J, 0, 0), cof3(J, 1, 0), cof3(J, 2, 0)};
        double det = c0[0] * J[0] + c0[1] * J[3] + c0[2] * J[6];
        double invdet = 1.0 / det;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) invJ[i * 3 + j] = cof3(J, j, i) * invdet;
        double d = det3_helper(J, 0, 1, 2) - det3_helper(J, 1, 0, 2) + det3_helper(J, 2, 0, 1);
        *measure = fabs(d) / 6;
        return 0;
    }
    if (M == 2 && N == 3) { /* simplex.h:190-191 manifold: (J^T J)^{-1} J^T, 0.5*|J0 x J1| */
        double a = 0, b = 0, d = 0;
        for (int r = 0; r < 3; ++r) { a += J[r * 2] * J[r * 2]; b += J[r * 2] * J[r * 2 + 1]; d += J[r * 2 + 1] * J[r * 2 + 1]; }
        double det = a * d - b * b, invdet = 1.0 / det;
        double g[4] = {d * invdet, -b * invdet, -b * invdet, a * invdet};
        for (int i = 0; i < 2; ++i)
            for (int r = 0; r < 3; ++r) invJ[i * 3 + r] = g[i * 2] * J[r * 2] + g[i * 2 + 1] * J[r * 2 + 1];
        double cx = J[1 * 2] * J[2 * 2 + 1] - J[2 * 2] * J[1 * 2 + 1];
        double cy = J[2 * 2] * J[0 * 2 + 1] - J[0 * 2] * J[2 * 2 + 1];
        double cz = J[0 * 2] * J[1 * 2 + 1] - J[1 * 2] * J[0 * 2 + 1];
        *measure = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
        return 0;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * weak forms and the (i,j) local integral
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int M, N, R, nb, nq;
    double coeff[ORC_MAXNB * ORC_MAXNB];
    double qn[ORC_MAXNQ * ORC_MAXD], qw[ORC_MAXNQ];
} orc_fe;

static int fe_init(orc_fe* fe, int M, int N, int R) {
    fe->M = M; fe->N = N; fe->R = R;
    fe->nb = orc_n_basis(M, R);
    fe->nq = orc_n_quad(M, R);
    if (fe->nb > ORC_MAXNB || fe->nq > ORC_MAXNQ) return -1;
    orc_ref_basis_coeffs(M, R, fe->coeff);
    return orc_quadrature(M, R, fe->qn, fe->qw);
}

/* g = (invJ^T) * grad psi (fem_assembler.h:81 stores invJ().transpose(); operators apply `invJ * nabla_psi`,
 * i.e. row r of invJ^T dotted with the gradient, accumulated from 0 upward: fields/dot_product.h:60-72) */
static void phys_grad(const orc_fe* fe, const double* invJ, const double* c, const double* p, double* g) {
    double gr[ORC_MAXD];
    for (int m = 0; m < fe->M; ++m) gr[m] = orc_poly_grad(fe->M, fe->R, c, m, p);
    for (int r = 0; r < fe->N; ++r) {
        double s = 0;
        for (int m = 0; m < fe->M; ++m) s += invJ[m * fe->N + r] * gr[m];
        g[r] = s;
    }
}

/* value of the whole operator tree at quadrature node q for the pair (i,j) of cell e:
 * operators/laplacian.h:37-44, diffusion.h:48-55, advection.h:49-56, reaction.h:47-53, dt.h:34-36,
 * combined left to right as differential_expressions.h:62-64 does. */
static double weak_form_at(const orc_fe* fe, const double* invJ, int i, int j, int q, int e, int n_terms,
                           const orc_term* terms) {
    const double* p = fe->qn + q * fe->M;
    const double* ci = fe->coeff + i * fe->nb;
    const double* cj = fe->coeff + j * fe->nb;
    int N = fe->N;
    double total = 0;
    for (int t = 0; t < n_terms; ++t) {
        const orc_term* T = &terms[t];
        double val = 0;
        switch (T->kind) {
        case ORC_LAPLACIAN: {
            double gi[ORC_MAXD], gj[ORC_MAXD], s = 0;
            phys_grad(fe, invJ, ci, p, gi);
            phys_grad(fe, invJ, cj, p, gj);
            for (int r = 0; r < N; ++r) s += gi[r] * gj[r];
            val = -s;
        } break;
        case ORC_DIFFUSION: {
            const double* K = T->coeff + (T->space_varying ? (size_t)(fe->nq * e + q) * N * N : 0);
            double gi[ORC_MAXD], gj[ORC_MAXD], kg[ORC_MAXD], s = 0;
            phys_grad(fe, invJ, ci, p, gi);
            phys_grad(fe, invJ, cj, p, gj);
            for (int r = 0; r < N; ++r) {
                double a = 0;
                for (int c = 0; c < N; ++c) a += K[c * N + r] * gj[c];
                kg[r] = a;
            }
            for (int r = 0; r < N; ++r) s += gi[r] * kg[r];
            val = -s;
        } break;
        case ORC_ADVECTION: {
            const double* b = T->coeff + (T->space_varying ? (size_t)(fe->nq * e + q) * N : 0);
            double gj[ORC_MAXD], s = 0;
            phys_grad(fe, invJ, cj, p, gj);
            for (int r = 0; r < N; ++r) s += gj[r] * b[r];
            val = orc_poly_eval(fe->M, fe->R, ci, p) * s;
        } break;
        case ORC_REACTION: {
            double c = T->coeff[T->space_varying ? (size_t)(fe->nq * e + q) : 0];
            val = c * orc_poly_eval(fe->M, fe->R, ci, p) * orc_poly_eval(fe->M, fe->R, cj, p);
        } break;
        default: val = 0; /* dT: zero field */
        }
        val = T->scale * val;
        total = (t == 0) ? val : total + val;
    }
    return total;
}

/* integrator.h:93-106 integrate_weak_form: (sum_q f(p_q) w_q) * measure */
static double integrate_weak_form(const orc_fe* fe, const double* invJ, double measure, int i, int j, int e,
                                  int n_terms, const orc_term* terms) {
    double value = 0;
    for (int q = 0; q < fe->nq; ++q) value += weak_form_at(fe, invJ, i, j, q, e, n_terms, terms) * fe->qw[q];
    return value * measure;
}

/* the nb x nb local matrix of one cell, as fem_operators_test.cpp:59-78 computes it. out[i*nb+j]. */
int orc_local_matrix(int M, int N, int R, const double* v, int cell_id, int n_terms, const orc_term* terms,
                     double* out) {
    orc_fe fe;
    double J[9], invJ[9], measure;
    if (fe_init(&fe, M, N, R)) return -1;
    if (orc_cell_geometry(M, N, v, J, invJ, &measure)) return -1;
    for (int i = 0; i < fe.nb; ++i)
        for (int j = 0; j < fe.nb; ++j)
            out[i * fe.nb + j] = integrate_weak_form(&fe, invJ, measure, i, j, cell_id, n_terms, terms);
    return 0;
}

/* physical gradients invJ^T grad psi_i at reference point p (lagrangian_basis_test.cpp:150-197) */
int orc_physical_gradients(int M, int N, int R, const double* v, const double* p, double* g /* nb*N */) {
    orc_fe fe;
    double J[9], invJ[9], measure;
    if (fe_init(&fe, M, N, R)) return -1;
    if (orc_cell_geometry(M, N, v, J, invJ, &measure)) return -1;
    for (int i = 0; i < fe.nb; ++i) phys_grad(&fe, invJ, fe.coeff + i * fe.nb, p, g + i * N);
    return 0;
}

static void cell_vertices(int M, int N, int n_nodes, const double* nodes, const int32_t* cells, int e, double* v) {
    for (int k = 0; k <= M; ++k)
        for (int r = 0; r < N; ++r) v[k * N + r] = nodes[(size_t)r * n_nodes + cells[(size_t)e * (M + 1) + k]];
}

/* ------------------------------------------------------------------------------------------
 * Eigen 3.4.0 SparseMatrix::setFromTriplets restated (column-major target):
 *   pass 1 count entries per row of the transposed temporary, pass 2 insert in triplet order,
 *   pass 3 collapse duplicates inside each row, adding later ones onto the first occurrence
 *   (left-to-right sum, nothing pruned), pass 4 transposed copy => sorted inner indices.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int32_t r, c; double v; } triplet;

static int64_t set_from_triplets(int n, const triplet* t, int64_t nt, int32_t** outer_o, int32_t** inner_o,
                                 double** val_o) {
    int64_t* rp = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    for (int64_t k = 0; k < nt; ++k) rp[t[k].r + 1]++;
    for (int i = 0; i < n; ++i) rp[i + 1] += rp[i];
    int32_t* cj = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nt ? nt : 1));
    double* cv = (double*)malloc(sizeof(double) * (size_t)(nt ? nt : 1));
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
    for (int i = 0; i < n; ++i) fill[i] = rp[i];
    for (int64_t k = 0; k < nt; ++k) {
        int64_t p = fill[t[k].r]++;
        cj[p] = t[k].c;
        cv[p] = t[k].v;
    }
    /* collapseDuplicates */
    int64_t* wi = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
    for (int i = 0; i < n; ++i) wi[i] = -1;
    int64_t* rp2 = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    int64_t count = 0;
    for (int i = 0; i < n; ++i) {
        int64_t start = count;
        for (int64_t k = rp[i]; k < rp[i + 1]; ++k) {
            int32_t c = cj[k];
            if (wi[c] >= start) {
                cv[wi[c]] += cv[k];
            } else {
                cv[count] = cv[k];
                cj[count] = c;
                wi[c] = count;
                ++count;
            }
        }
        rp2[i] = start;
    }
    rp2[n] = count;
    /* transposed copy: row-major temporary -> column-major result, inner (row) indices ascending */
    int32_t* outer = (int32_t*)calloc((size_t)n + 1, sizeof(int32_t));
    int32_t* inner = (int32_t*)malloc(sizeof(int32_t) * (size_t)(count ? count : 1));
    double* val = (double*)malloc(sizeof(double) * (size_t)(count ? count : 1));
    for (int64_t k = 0; k < count; ++k) outer[cj[k] + 1]++;
    for (int i = 0; i < n; ++i) outer[i + 1] += outer[i];
    for (int i = 0; i < n; ++i) fill[i] = outer[i];
    for (int i = 0; i < n; ++i)
        for (int64_t k = rp2[i]; k < rp2[i + 1]; ++k) {
            int64_t p = fill[cj[k]]++;
            inner[p] = i;
            val[p] = cv[k];
        }
    free(rp); free(cj); free(cv); free(fill); free(wi); free(rp2);
    *outer_o = outer; *inner_o = inner; *val_o = val;
    return count;
}

/* Eigen 3.4.0 SparseSelfAdjointView<Lower> -> SparseMatrix assignment (permute_symm_to_fullsymm):
 * every stored (i,j), i>j, is mirrored to (j,i); diagonal kept once; result columns have ascending rows. */
static int64_t selfadjoint_lower_to_full(int n, const int32_t* outer, const int32_t* inner, const double* val,
                                         int32_t** outer_o, int32_t** inner_o, double** val_o) {
    int64_t* cnt = (int64_t*)calloc((size_t)n + 1, sizeof(int64_t));
    for (int j = 0; j < n; ++j)
        for (int32_t k = outer[j]; k < outer[j + 1]; ++k) {
            int i = inner[k];
            if (i == j) cnt[j + 1]++;
            else if (i > j) { cnt[j + 1]++; cnt[i + 1]++; }
        }
    for (int j = 0; j < n; ++j) cnt[j + 1] += cnt[j];
    int64_t nnz = cnt[n];
    int32_t* o = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
    int32_t* in = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz ? nnz : 1));
    double* v = (double*)malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
    int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n ? n : 1));
    for (int j = 0; j <= n; ++j) o[j] = (int32_t)cnt[j];
    for (int j = 0; j < n; ++j) fill[j] = cnt[j];
    for (int j = 0; j < n; ++j)
        for (int32_t k = outer[j]; k < outer[j + 1]; ++k) {
            int i = inner[k];
            if (i == j) { int64_t p = fill[j]++; in[p] = i; v[p] = val[k]; }
            else if (i > j) {
                int64_t p = fill[j]++; in[p] = i; v[p] = val[k];
                p = fill[i]++; in[p] = j; v[p] = val[k];
            }
        }
    free(cnt); free(fill);
    *outer_o = o; *inner_o = in; *val_o = v;
    return nnz;
}

void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * Assembler<FEM,...>::discretize_operator  (finite_elements/fem_assembler.h:52-121)
 * returns nnz; CSC arrays (int32 outer[n_dofs+1], inner[nnz], double val[nnz]) malloc'ed.
 * ---------------------------------------------------------------------------------------- */
int64_t orc_assemble_operator(int M, int N, int R, int n_nodes, int n_cells, const double* nodes,
                              const int32_t* cells, int n_dofs, const int32_t* dofs, int n_terms,
                              const orc_term* terms, int symmetric, int32_t** outer, int32_t** inner,
                              double** val) {
    orc_fe fe;
    if (fe_init(&fe, M, N, R)) return -1;
    int nb = fe.nb;
    size_t cap = (size_t)n_cells * (size_t)(symmetric ? nb * (nb + 1) / 2 + nb : nb * nb) + 1;
    triplet* tl = (triplet*)malloc(sizeof(triplet) * cap);
    int64_t nt = 0;
    double v[(ORC_MAXD + 1) * ORC_MAXD], J[9], invJ[9], measure;
    for (int e = 0; e < n_cells; ++e) { /* fem_assembler.h:79 */
        cell_vertices(M, N, n_nodes, nodes, cells, e, v);
        orc_cell_geometry(M, N, v, J, invJ, &measure);
        for (int i = 0; i < nb; ++i) {
            int32_t di = dofs[(size_t)i * n_cells + e];
            for (int j = 0; j < nb; ++j) {
                int32_t dj = dofs[(size_t)j * n_cells + e];
                if (symmetric && !(di >= dj)) continue; /* :94-96 */
                if ((size_t)nt >= cap) { cap *= 2; tl = (triplet*)realloc(tl, sizeof(triplet) * cap); }
                tl[nt].r = di;
                tl[nt].c = dj;
                tl[nt].v = integrate_weak_form(&fe, invJ, measure, i, j, e, n_terms, terms);
                ++nt;
            }
        }
    }
    int32_t *o, *in;
    double* vv;
    int64_t nnz = set_from_triplets(n_dofs, tl, nt, &o, &in, &vv); /* :112-113 */
    free(tl);
    if (symmetric) { /* :116-117 selfadjointView<Lower>() */
        int32_t *o2, *in2;
        double* v2;
        nnz = selfadjoint_lower_to_full(n_dofs, o, in, vv, &o2, &in2, &v2);
        free(o); free(in); free(vv);
        o = o2; in = in2; vv = v2;
    }
    *outer = o; *inner = in; *val = vv;
    return nnz;
}

/* Best-effort all-core variant of the same algorithm (SURVEY 8d "best-effort CPU" line; bench.py --impl reference).
 * The reference itself is strictly serial.  Every cell emits its triplets into a precomputed slot range (cells
 * ascending, i outer, j inner), so the triplet list -- and therefore the matrix -- is bit-identical to the serial
 * routine for any number of threads; setFromTriplets and the mirror pass stay serial.  Without OpenMP this is the
 * serial routine. */
int64_t orc_assemble_operator_mt(int M, int N, int R, int n_nodes, int n_cells, const double* nodes,
                                 const int32_t* cells, int n_dofs, const int32_t* dofs, int n_terms,
                                 const orc_term* terms, int symmetric, int n_threads, int32_t** outer, int32_t** inner,
                                 double** val) {
#ifndef _OPENMP
    (void)n_threads;
    return orc_assemble_operator(M, N, R, n_nodes, n_cells, nodes, cells, n_dofs, dofs, n_terms, terms, symmetric, outer,
                                 inner, val);
#else
    orc_fe fe;
    if (fe_init(&fe, M, N, R)) return -1;
    const int nb = fe.nb;
    int64_t* off = (int64_t*)malloc(sizeof(int64_t) * ((size_t)n_cells + 1));
    off[0] = 0;
    for (int e = 0; e < n_cells; ++e) { /* triplets emitted by cell e */
        int cnt = 0;
        for (int i = 0; i < nb; ++i)
            for (int j = 0; j < nb; ++j)
                cnt += !(symmetric && !(dofs[(size_t)i * n_cells + e] >= dofs[(size_t)j * n_cells + e]));
        off[e + 1] = off[e] + cnt;
    }
    const int64_t nt = off[n_cells];
    triplet* tl = (triplet*)malloc(sizeof(triplet) * (size_t)(nt ? nt : 1));
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(static) num_threads(n_threads)
    for (int e = 0; e < n_cells; ++e) {
        double v[(ORC_MAXD + 1) * ORC_MAXD], J[9], invJ[9], measure;
        cell_vertices(M, N, n_nodes, nodes, cells, e, v);
        orc_cell_geometry(M, N, v, J, invJ, &measure);
        int64_t k = off[e];
        for (int i = 0; i < nb; ++i) {
            int32_t di = dofs[(size_t)i * n_cells + e];
            for (int j = 0; j < nb; ++j) {
                int32_t dj = dofs[(size_t)j * n_cells + e];
                if (symmetric && !(di >= dj)) continue;
                tl[k].r = di;
                tl[k].c = dj;
                tl[k].v = integrate_weak_form(&fe, invJ, measure, i, j, e, n_terms, terms);
                ++k;
            }
        }
    }
    free(off);
    int32_t *o, *in;
    double* vv;
    int64_t nnz = set_from_triplets(n_dofs, tl, nt, &o, &in, &vv);
    free(tl);
    if (symmetric) {
        int32_t *o2, *in2;
        double* v2;
        nnz = selfadjoint_lower_to_full(n_dofs, o, in, vv, &o2, &in2, &v2);
        free(o); free(in); free(vv);
        o = o2; in = in2; vv = v2;
    }
    *outer = o; *inner = in; *val = vv;
    return nnz;
#endif
}

/* Assembler::discretize_forcing (fem_assembler.h:122-136) with matrix-of-values forcing
 * (integrator.h:74-90, fallback branch): f[nq*e+q]. */
int orc_assemble_forcing(int M, int N, int R, int n_nodes, int n_cells, const double* nodes, const int32_t* cells,
                         int n_dofs, const int32_t* dofs, const double* f_quad, double* b) {
    orc_fe fe;
    if (fe_init(&fe, M, N, R)) return -1;
    double v[(ORC_MAXD + 1) * ORC_MAXD], J[9], invJ[9], measure;
    for (int i = 0; i < n_dofs; ++i) b[i] = 0;
    for (int e = 0; e < n_cells; ++e) {
        cell_vertices(M, N, n_nodes, nodes, cells, e, v);
        orc_cell_geometry(M, N, v, J, invJ, &measure);
        for (int i = 0; i < fe.nb; ++i) {
            double value = 0;
            for (int q = 0; q < fe.nq; ++q) {
                double phi = orc_poly_eval(M, R, fe.coeff + i * fe.nb, fe.qn + q * M);
                value += (f_quad[(size_t)fe.nq * e + q] * phi) * fe.qw[q];
            }
            b[dofs[(size_t)i * n_cells + e]] += value * measure;
        }
    }
    return 0;
}

/* Integrator::quadrature_nodes (integrator.h:109-121): row nq*e+q = J p_q + v0; out column-major (n_cells*nq) x N */
int orc_quadrature_nodes(int M, int N, int R, int n_nodes, int n_cells, const double* nodes, const int32_t* cells,
                         double* out) {
    orc_fe fe;
    if (fe_init(&fe, M, N, R)) return -1;
    double v[(ORC_MAXD + 1) * ORC_MAXD], J[9], invJ[9], measure;
    size_t rows = (size_t)n_cells * fe.nq;
    for (int e = 0; e < n_cells; ++e) {
        cell_vertices(M, N, n_nodes, nodes, cells, e, v);
        orc_cell_geometry(M, N, v, J, invJ, &measure);
        for (int q = 0; q < fe.nq; ++q)
            for (int r = 0; r < N; ++r) {
                double s = 0;
                for (int m = 0; m < M; ++m) s += J[r * M + m] * fe.qn[q * M + m];
                out[(size_t)r * rows + (size_t)fe.nq * e + q] = s + v[r];
            }
    }
    return 0;
}

/* Integrator::integrate(mesh, f) with f == 1 (integration_test.cpp:46-58,79): sum of (sum_q w_q) * measure */
double orc_integrate_one(int M, int N, int R, int n_nodes, int n_cells, const double* nodes, const int32_t* cells) {
    orc_fe fe;
    if (fe_init(&fe, M, N, R)) return NAN;
    double v[(ORC_MAXD + 1) * ORC_MAXD], J[9], invJ[9], measure, total = 0;
    for (int e = 0; e < n_cells; ++e) {
        cell_vertices(M, N, n_nodes, nodes, cells, e, v);
        orc_cell_geometry(M, N, v, J, invJ, &measure);
        double value = 0;
        for (int q = 0; q < fe.nq; ++q) value += 1.0 * fe.qw[q];
        total += value * measure;
    }
    return total;
}

/* ------------------------------------------------------------------------------------------
 * mesh topology needed for P2 dofs
 *   2D: Triangulation<2,N> ctor (geometry/triangulation.h:143-196): edge id = first occurrence scanning
 *       cells ascending x local pairs (0,1),(0,2),(1,2) (utils/combinatorics.h:37-51); boundary edge <=>
 *       seen by exactly one cell.
 *   3D: Triangulation<3,3> ctor (:319-399): faces (0,1,2),(0,1,3),(0,2,3),(1,2,3) of each cell; edges are
 *       numbered inside each NEW face from its sorted node triple, pairs (0,1),(0,2),(1,2); boundary edge
 *       <=> both nodes on the boundary (:376).
 * A sort-based formulation (key = node pair, value = scan position, keep the minimum, rank by it)
 * reproduces the hash-map scan exactly.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int32_t a, b; int64_t pos; int32_t id; } edge_rec;
static int cmp_edge_key(const void* x, const void* y) {
    const edge_rec *p = (const edge_rec*)x, *q = (const edge_rec*)y;
    if (p->a != q->a) return p->a < q->a ? -1 : 1;
    if (p->b != q->b) return p->b < q->b ? -1 : 1;
    return p->pos < q->pos ? -1 : (p->pos > q->pos ? 1 : 0);
}
static int cmp_edge_pos(const void* x, const void* y) {
    const edge_rec *p = (const edge_rec*)x, *q = (const edge_rec*)y;
    return p->pos < q->pos ? -1 : (p->pos > q->pos ? 1 : 0);
}
static void sort3(int32_t* f) {
    int32_t t;
    if (f[0] > f[1]) { t = f[0]; f[0] = f[1]; f[1] = t; }
    if (f[1] > f[2]) { t = f[1]; f[1] = f[2]; f[2] = t; }
    if (f[0] > f[1]) { t = f[0]; f[0] = f[1]; f[1] = t; }
}

/* ------------------------------------------------------------------------------------------
 * Full mesh topology (next-row N3): a LITERAL restatement of the hash-map scans of
 *   Triangulation<2,N>::Triangulation   geometry/triangulation.h:143-196
 *   Triangulation<3,3>::Triangulation   geometry/triangulation.h:319-399
 * Facets (edges of triangles, faces of tetrahedra) are visited cells ascending x local patterns of
 * combinations<M, M+1>() (utils/combinatorics.h:37-51); a facet met for the first time gets the next id and waits in the
 * map with its cell; met again it links the two cells as neighbours -- neighbors(cell, j) with j the first local vertex
 * not on the facet (:156-166, :334-344) --, fills facet_to_cells(.,1), clears its boundary marker and LEAVES the map.
 * 3D: the edges are numbered inside every NEW face from its sorted node triple, pairs (0,1),(0,2),(1,2) (:356-373);
 * an edge is on the boundary iff both end nodes are (:370); edge_to_cells collects every cell of a face that holds the
 * edge (a set: returned here as ascending lists).
 * Open addressing with tombstones stands in for std::unordered_map (only find / emplace / erase are used).
 * ---------------------------------------------------------------------------------------- */
typedef struct { int32_t k[3]; int32_t id, cell; int state; /* 0 empty, 1 live, 2 erased */ } topo_slot;
static uint64_t topo_hash(const int32_t* k, int len) {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < len; ++i) { h ^= (uint64_t)(uint32_t)k[i]; h *= 1099511628211ull; h ^= h >> 29; }
    return h;
}
static topo_slot* topo_find(topo_slot* tab, size_t cap, const int32_t* k, int len, int for_insert) {
    size_t i = (size_t)(topo_hash(k, len) % cap);
    topo_slot* grave = NULL;
    for (;;) {
        topo_slot* s = &tab[i];
        if (s->state == 0) return for_insert ? (grave ? grave : s) : NULL;
        if (s->state == 2) { if (!grave) grave = s; }
        else {
            int eq = 1;
            for (int t = 0; t < len; ++t) eq &= (s->k[t] == k[t]);
            if (eq) return s;
        }
        i = (i + 1 == cap) ? 0 : i + 1;
    }
}
static void sort_small(int32_t* v, int len) {
    for (int i = 1; i < len; ++i) { int32_t x = v[i]; int j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; } v[j + 1] = x; }
}
static int cmp_i32(const void* a, const void* b) { int32_t x = *(const int32_t*)a, y = *(const int32_t*)b; return (x > y) - (x < y); }

/* Sizes first (capacity of the outputs): n_facets <= n_cells * (M + 1); 3D n_edges <= 3 * n_facets.
 * Outputs (row-major, caller allocated at capacity):
 *   neighbors      n_cells x (M+1)      (-1: no neighbour across the facet opposite to that vertex)
 *   facets         cap x M              sorted node ids      (2D: edges(), 3D: faces())
 *   cell_to_facets n_cells x (M+1)      (2D: cell_to_edges(), 3D: cell_to_faces())
 *   facet_to_cells cap x 2              (first cell, second cell or -1)
 *   facet_boundary cap
 *   3D only: edges cap3 x 2, face_to_edges cap x 3, edge_boundary cap3, edge_cell_ptr cap3 + 1, edge_cells <= 6 n_cells
 * Returns n_facets; *n_edges_out = number of edges (2D: == n_facets). */
int orc_mesh_topology(int M, int n_cells, const int32_t* cells, const uint8_t* boundary_nodes, int32_t* neighbors,
                      int32_t* facets, int32_t* cell_to_facets, int32_t* facet_to_cells, uint8_t* facet_boundary,
                      int32_t* edges, int32_t* face_to_edges, uint8_t* edge_boundary, int32_t* edge_cell_ptr,
                      int32_t* edge_cells, int* n_edges_out) {
    const int nv = M + 1, fl = M;   /* nodes per facet */
    static const int pat2[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    static const int pat3[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
    const size_t cap = (size_t)n_cells * nv * 2 + 16;
    topo_slot* fmap = (topo_slot*)calloc(cap, sizeof(topo_slot));
    const size_t ecap = (M == 3) ? (size_t)n_cells * 6 * 2 + 16 : 1;
    topo_slot* emap = (topo_slot*)calloc(ecap, sizeof(topo_slot));
    /* 3D: (edge, cell) bindings in insertion order; made unique + ascending at the end (an unordered_set in the reference) */
    size_t nbind = 0, bind_cap = (M == 3) ? (size_t)n_cells * 12 + 16 : 1;
    int32_t* bind_e = (int32_t*)malloc(sizeof(int32_t) * bind_cap);
    int32_t* bind_c = (int32_t*)malloc(sizeof(int32_t) * bind_cap);
    for (size_t t = 0; t < (size_t)n_cells * nv; ++t) neighbors[t] = -1;
    int facet_id = 0, edge_id = 0;
    for (int i = 0; i < n_cells; ++i) {
        const int32_t* c = cells + (size_t)i * nv;
        for (int j = 0; j < nv; ++j) {
            int32_t f[3];
            for (int k = 0; k < fl; ++k) f[k] = c[M == 2 ? pat2[j][k] : pat3[j][k]];
            sort_small(f, fl);
            topo_slot* it = topo_find(fmap, cap, f, fl, 0);
            if (!it) {   /* never processed facet */
                for (int k = 0; k < fl; ++k) facets[(size_t)facet_id * fl + k] = f[k];
                facet_to_cells[2 * (size_t)facet_id] = i;
                facet_to_cells[2 * (size_t)facet_id + 1] = -1;
                facet_boundary[facet_id] = 1;
                topo_slot* s = topo_find(fmap, cap, f, fl, 1);
                for (int k = 0; k < fl; ++k) s->k[k] = f[k];
                s->id = facet_id; s->cell = i; s->state = 1;
                cell_to_facets[(size_t)i * nv + j] = facet_id;
                if (M == 3) {   /* ids of the edges of the new face */
                    for (int k = 0; k < 3; ++k) {
                        int32_t e[2] = {f[pat2[k][0]], f[pat2[k][1]]};
                        sort_small(e, 2);
                        topo_slot* ie = topo_find(emap, ecap, e, 2, 0);
                        int id;
                        if (!ie) {
                            edges[2 * (size_t)edge_id] = e[0]; edges[2 * (size_t)edge_id + 1] = e[1];
                            topo_slot* se = topo_find(emap, ecap, e, 2, 1);
                            se->k[0] = e[0]; se->k[1] = e[1]; se->id = edge_id; se->state = 1;
                            edge_boundary[edge_id] = boundary_nodes ? (boundary_nodes[e[0]] && boundary_nodes[e[1]]) : 0;
                            id = edge_id++;
                        } else id = ie->id;
                        face_to_edges[3 * (size_t)facet_id + k] = id;
                        bind_e[nbind] = id; bind_c[nbind] = i; ++nbind;
                    }
                }
                ++facet_id;
            } else {
                const int h = it->id, kc = it->cell;
                /* first local vertex of the cell that is not a node of the facet */
                for (int side = 0; side < 2; ++side) {
                    const int cell = side ? i : kc, other = side ? kc : i;
                    const int32_t* cc = cells + (size_t)cell * nv;
                    int jj = 0;
                    for (; jj < nv; ++jj) {
                        int found = 0;
                        for (int k = 0; k < fl; ++k) found |= (facets[(size_t)h * fl + k] == cc[jj]);
                        if (!found) break;
                    }
                    neighbors[(size_t)cell * nv + jj] = other;
                }
                if (M == 3)
                    for (int k = 0; k < 3; ++k) { bind_e[nbind] = face_to_edges[3 * (size_t)h + k]; bind_c[nbind] = i; ++nbind; }
                cell_to_facets[(size_t)i * nv + j] = h;
                facet_to_cells[2 * (size_t)h + 1] = i;
                facet_boundary[h] = 0;
                it->state = 2;   /* erase */
            }
        }
    }
    if (M == 3) {   /* edge -> cells as ascending unique lists */
        int32_t* cnt = (int32_t*)calloc((size_t)edge_id + 1, sizeof(int32_t));
        for (size_t t = 0; t < nbind; ++t) cnt[bind_e[t]]++;
        int32_t* off = (int32_t*)malloc(sizeof(int32_t) * ((size_t)edge_id + 1));
        int32_t acc = 0;
        for (int e = 0; e < edge_id; ++e) { off[e] = acc; acc += cnt[e]; cnt[e] = 0; }
        off[edge_id] = acc;
        int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * (nbind ? nbind : 1));
        for (size_t t = 0; t < nbind; ++t) tmp[off[bind_e[t]] + cnt[bind_e[t]]++] = bind_c[t];
        int32_t w = 0;
        for (int e = 0; e < edge_id; ++e) {
            qsort(tmp + off[e], cnt[e], sizeof(int32_t), cmp_i32);
            edge_cell_ptr[e] = w;
            for (int t = 0; t < cnt[e]; ++t)
                if (t == 0 || tmp[off[e] + t] != tmp[off[e] + t - 1]) edge_cells[w++] = tmp[off[e] + t];
        }
        edge_cell_ptr[edge_id] = w;
        free(cnt); free(off); free(tmp);
    }
    free(fmap); free(emap); free(bind_e); free(bind_c);
    *n_edges_out = (M == 3) ? edge_id : facet_id;
    return facet_id;
}

/* Enumerate mesh edges.  cell_edges: n_cells x ne (ne = 3 in 2D, 6 in 3D, row-major) gives, for every
 * local vertex pair in the order pairs2[] / pairs3[] below, the global edge id.  edges_out: n_edges x 2
 * sorted node pairs.  edge_boundary: 0/1.  Returns n_edges. */
static const int pairs2[3][2] = {{0, 1}, {0, 2}, {1, 2}};
static const int pairs3[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
static const int faces3[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};

int orc_enumerate_edges(int M, int n_cells, const int32_t* cells, const uint8_t* boundary_nodes,
                        int32_t* cell_edges, int32_t** edges_out, uint8_t** edge_boundary_out) {
    int per_cell = (M == 2) ? 3 : 12; /* scan slots per cell: 3 pairs, or 4 faces x 3 pairs */
    size_t ns = (size_t)n_cells * per_cell;
    edge_rec* rec = (edge_rec*)malloc(sizeof(edge_rec) * (ns ? ns : 1));
    for (int e = 0; e < n_cells; ++e) {
        const int32_t* c = cells + (size_t)e * (M + 1);
        if (M == 2) {
            for (int j = 0; j < 3; ++j) {
                int32_t a = c[pairs2[j][0]], b = c[pairs2[j][1]];
                edge_rec* r = &rec[(size_t)e * 3 + j];
                r->a = a < b ? a : b; r->b = a < b ? b : a; r->pos = (int64_t)e * 3 + j;
            }
        } else {
            for (int f = 0; f < 4; ++f) {
                int32_t face[3] = {c[faces3[f][0]], c[faces3[f][1]], c[faces3[f][2]]};
                sort3(face);
                for (int k = 0; k < 3; ++k) {
                    edge_rec* r = &rec[(size_t)e * 12 + f * 3 + k];
                    r->a = face[pairs2[k][0]]; r->b = face[pairs2[k][1]]; r->pos = (int64_t)e * 12 + f * 3 + k;
                }
            }
        }
    }
    qsort(rec, ns, sizeof(edge_rec), cmp_edge_key);
    /* unique keys, each represented by its first scan position; count multiplicity for the 2D boundary flag */
    size_t nu = 0;
    int32_t* mult = (int32_t*)malloc(sizeof(int32_t) * (ns ? ns : 1));
    edge_rec* uq = (edge_rec*)malloc(sizeof(edge_rec) * (ns ? ns : 1));
    for (size_t k = 0; k < ns; ++k) {
        if (k == 0 || rec[k].a != rec[k - 1].a || rec[k].b != rec[k - 1].b) { uq[nu] = rec[k]; mult[nu] = 1; uq[nu].id = (int32_t)nu; ++nu; }
        else mult[nu - 1]++;
    }
    /* rank unique edges by first scan position => reference edge ids */
    edge_rec* byp = (edge_rec*)malloc(sizeof(edge_rec) * (nu ? nu : 1));
    memcpy(byp, uq, sizeof(edge_rec) * nu);
    qsort(byp, nu, sizeof(edge_rec), cmp_edge_pos);
    int32_t* key_to_id = (int32_t*)malloc(sizeof(int32_t) * (nu ? nu : 1));
    int32_t* edges = (int32_t*)malloc(sizeof(int32_t) * 2 * (nu ? nu : 1));
    uint8_t* eb = (uint8_t*)malloc(nu ? nu : 1);
    for (size_t r = 0; r < nu; ++r) {
        key_to_id[byp[r].id] = (int32_t)r;
        edges[2 * r] = byp[r].a; edges[2 * r + 1] = byp[r].b;
        if (M == 2) eb[r] = (mult[byp[r].id] == 1);
        else eb[r] = boundary_nodes ? (boundary_nodes[byp[r].a] && boundary_nodes[byp[r].b]) : 0;
    }
    /* per-cell lookup by binary search in the key-sorted unique list */
    int ne = (M == 2) ? 3 : 6;
    for (int e = 0; e < n_cells; ++e) {
        const int32_t* c = cells + (size_t)e * (M + 1);
        for (int j = 0; j < ne; ++j) {
            int32_t a = (M == 2) ? c[pairs2[j][0]] : c[pairs3[j][0]];
            int32_t b = (M == 2) ? c[pairs2[j][1]] : c[pairs3[j][1]];
            if (a > b) { int32_t t = a; a = b; b = t; }
            size_t lo = 0, hi = nu;
            while (lo < hi) {
                size_t mid = (lo + hi) / 2;
                if (uq[mid].a < a || (uq[mid].a == a && uq[mid].b < b)) lo = mid + 1; else hi = mid;
            }
            cell_edges[(size_t)e * ne + j] = key_to_id[lo];
        }
    }
    free(rec); free(mult); free(uq); free(byp); free(key_to_id);
    *edges_out = edges; *edge_boundary_out = eb;
    return (int)nu;
}

/* LagrangianBasis::enumerate_dofs (basis/lagrangian_basis.h:94-136).
 * R == 1: dofs = cells, boundary = node markers.  R == 2 on triangles: dof n_nodes + edge_id at local slot
 * 3 + pair index.  R == 2 on tetrahedra is NOT in the reference (it does not compile there):
 * extension A10 -- slot = the reference node of ReferenceElement<3,2> (reference_element.h:93-96) that is the
 * midpoint of the vertex pair: 4<->(1,2) 5<->(0,2) 6<->(0,1) 7<->(1,3) 8<->(2,3) 9<->(0,3).
 * dofs_out column-major n_cells x nb; boundary_dofs_out must hold n_nodes + n_edges bytes. Returns n_dofs. */
int orc_enumerate_dofs(int M, int R, int n_nodes, int n_cells, const int32_t* cells, const uint8_t* boundary_nodes,
                       int32_t* dofs_out, uint8_t* boundary_dofs_out) {
    int nb = orc_n_basis(M, R);
    for (int k = 0; k <= M; ++k)
        for (int e = 0; e < n_cells; ++e) dofs_out[(size_t)k * n_cells + e] = cells[(size_t)e * (M + 1) + k];
    for (int i = 0; i < n_nodes; ++i) boundary_dofs_out[i] = boundary_nodes ? boundary_nodes[i] : 0;
    if (R == 1) return n_nodes;
    int ne = (M == 2) ? 3 : 6;
    int32_t* ce = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_cells * ne);
    int32_t* edges;
    uint8_t* eb;
    int n_edges = orc_enumerate_edges(M, n_cells, cells, boundary_nodes, ce, &edges, &eb);
    /* slot of pair j: 2D = 3 + j; 3D: pairs3 order (0,1),(0,2),(0,3),(1,2),(1,3),(2,3) -> 6,5,9,4,7,8 */
    static const int slot3[6] = {6, 5, 9, 4, 7, 8};
    for (int e = 0; e < n_cells; ++e)
        for (int j = 0; j < ne; ++j) {
            int slot = (M == 2) ? 3 + j : slot3[j];
            dofs_out[(size_t)slot * n_cells + e] = n_nodes + ce[(size_t)e * ne + j];
        }
    for (int k = 0; k < n_edges; ++k) boundary_dofs_out[n_nodes + k] = eb[k];
    (void)nb;
    free(ce); free(edges); free(eb);
    return n_nodes + n_edges;
}

/* LagrangianBasis::dofs_coords (lagrangian_basis.h:159-183): vertices, then edge midpoints J*ref + v0 taken
 * from the first cell (ascending id) that holds the dof.  out column-major n_dofs x N. */
int orc_dofs_coords(int M, int N, int R, int n_nodes, int n_cells, const double* nodes, const int32_t* cells,
                    int n_dofs, const int32_t* dofs, double* out) {
    int nb = orc_n_basis(M, R);
    double ref[ORC_MAXNB * ORC_MAXD];
    orc_reference_nodes(M, R, ref);
    for (int r = 0; r < N; ++r)
        for (int i = 0; i < n_nodes; ++i) out[(size_t)r * n_dofs + i] = nodes[(size_t)r * n_nodes + i];
    if (R == 1) return 0;
    uint8_t* visited = (uint8_t*)calloc((size_t)n_dofs, 1);
    double v[(ORC_MAXD + 1) * ORC_MAXD], J[9], invJ[9], measure;
    for (int e = 0; e < n_cells; ++e) {
        cell_vertices(M, N, n_nodes, nodes, cells, e, v);
        orc_cell_geometry(M, N, v, J, invJ, &measure);
        for (int j = M + 1; j < nb; ++j) {
            int32_t d = dofs[(size_t)j * n_cells + e];
            if (visited[d]) continue;
            for (int r = 0; r < N; ++r) {
                double s = 0;
                for (int m = 0; m < M; ++m) s += J[r * M + m] * ref[j * M + m];
                out[(size_t)r * n_dofs + d] = s + v[r];
            }
            visited[d] = 1;
        }
    }
    free(visited);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Next-row N1: basis evaluation matrices Psi (basis/lagrangian_basis.h:203-283) and point location.
 * ---------------------------------------------------------------------------------------- */
/* Simplex::contains (geometry/simplex.h:115-128): barycentric coordinates z, OUTSIDE iff any z < -machine_epsilon
 * with machine_epsilon = 10 * DBL_EPSILON (utils/symbols.h:164); manifold cells add the supporting-plane test. */
static int cell_contains(int M, int N, const double* v, const double* invJ, const double* x) {
    const double meps = 10 * 2.220446049250313e-16;
    double z[ORC_MAXD + 1], d[ORC_MAXD], sum = 0;
    for (int r = 0; r < N; ++r) d[r] = x[r] - v[r];
    if (M == 2 && N == 3) {
        /* manifold cells (simplex.h:116-118): the point must lie on the supporting plane, distance =
         * |x - (B B^T (x - p) + p)| with the orthonormal basis of HyperPlane<2,3> (hyperplane.h:56-62,92-99) */
        double b0[3], b1[3], w[3], n0 = 0, n1 = 0, wb = 0, bb = 0;
        for (int r = 0; r < 3; ++r) { b0[r] = v[3 + r] - v[r]; n0 += b0[r] * b0[r]; }
        n0 = sqrt(n0);
        for (int r = 0; r < 3; ++r) b0[r] /= n0;
        for (int r = 0; r < 3; ++r) { w[r] = v[6 + r] - v[r]; wb += w[r] * b0[r]; bb += b0[r] * b0[r]; }
        for (int r = 0; r < 3; ++r) { b1[r] = w[r] - wb / bb * b0[r]; n1 += b1[r] * b1[r]; }
        n1 = sqrt(n1);
        for (int r = 0; r < 3; ++r) b1[r] /= n1;
        double c0 = 0, c1 = 0, dist = 0;
        for (int r = 0; r < 3; ++r) { c0 += b0[r] * d[r]; c1 += b1[r] * d[r]; }
        for (int r = 0; r < 3; ++r) {
            double pr = (b0[r] * c0 + b1[r] * c1) + v[r];
            dist += (x[r] - pr) * (x[r] - pr);
        }
        if (sqrt(dist) > meps) return 0;
    }
    for (int m = 0; m < M; ++m) {
        double t = 0;
        for (int r = 0; r < N; ++r) t += invJ[m * N + r] * d[r];
        z[m + 1] = t;
        sum += t;
    }
    z[0] = 1 - sum;
    for (int m = 0; m <= M; ++m)
        if (z[m] < -meps) return 0;
    return 1;
}

/* Triangulation::locate (triangulation.h:252-255, tree_search.h:71-86): id of a cell containing the point, -1 if none.
 * The reference walks the candidates of a KD-tree range query in std::unordered_set order, i.e. for a point shared by
 * several cells (on an edge or a vertex) the winner is implementation defined; this restatement (and the CUDA path)
 * pins it to the SMALLEST cell id.  Psi itself does not depend on the choice (the basis is continuous) except for
 * which explicit zeros are stored.  locs column-major n_locs x N. */
int orc_locate(int M, int N, int n_nodes, int n_cells, const double* nodes, const int32_t* cells, int n_locs,
               const double* locs, int32_t* ids) {
    if (M != N && !(M == 2 && N == 3)) return -1;
    double v[(ORC_MAXD + 1) * ORC_MAXD], J[9], invJ[9], measure, x[ORC_MAXD];
    for (int i = 0; i < n_locs; ++i) ids[i] = -1;
    for (int e = 0; e < n_cells; ++e) {
        cell_vertices(M, N, n_nodes, nodes, cells, e, v);
        orc_cell_geometry(M, N, v, J, invJ, &measure);
        for (int i = 0; i < n_locs; ++i) {
            if (ids[i] >= 0) continue;
            for (int r = 0; r < N; ++r) x[r] = locs[(size_t)r * n_locs + i];
            if (cell_contains(M, N, v, invJ, x)) ids[i] = e;
        }
    }
    return 0;
}

/* pointwise_evaluation::eval (lagrangian_basis.h:203-235): row i holds psi_h(invJ (p_i - v0)) at column dofs(e, h)
 * for the cell e containing p_i; rows of points outside the domain are empty.  Output = the triplet list in emission
 * order, n_basis slots per point (cols -1 / vals 0 for an outside point); ids[i] = cell of point i. */
int orc_eval_pointwise(int M, int N, int R, int n_nodes, int n_cells, const double* nodes, const int32_t* cells,
                       const int32_t* dofs, int n_locs, const double* locs, int32_t* ids, int32_t* cols, double* vals) {
    orc_fe fe;
    if (fe_init(&fe, M, N, R)) return -1;
    if (orc_locate(M, N, n_nodes, n_cells, nodes, cells, n_locs, locs, ids)) return -1;
    double v[(ORC_MAXD + 1) * ORC_MAXD], J[9], invJ[9], measure, xi[ORC_MAXD];
    for (int i = 0; i < n_locs; ++i) {
        int e = ids[i];
        for (int h = 0; h < fe.nb; ++h) { cols[(size_t)i * fe.nb + h] = -1; vals[(size_t)i * fe.nb + h] = 0; }
        if (e < 0) continue;
        cell_vertices(M, N, n_nodes, nodes, cells, e, v);
        orc_cell_geometry(M, N, v, J, invJ, &measure);
        for (int m = 0; m < M; ++m) {
            double t = 0;
            for (int r = 0; r < N; ++r) t += invJ[m * N + r] * (locs[(size_t)r * n_locs + i] - v[r]);
            xi[m] = t;
        }
        for (int h = 0; h < fe.nb; ++h) {
            cols[(size_t)i * fe.nb + h] = dofs[(size_t)h * n_cells + e];
            vals[(size_t)i * fe.nb + h] = orc_poly_eval(M, R, fe.coeff + h * fe.nb, xi);
        }
    }
    return 0;
}

/* areal_evaluation::eval (lagrangian_basis.h:238-283): incidence column-major n_sub x n_cells (entry == 1: the cell
 * belongs to the subdomain).  For subdomain k, cells ascending, h ascending: triplet (k, dofs(e,h), int_e psi_h / D_k)
 * with int_e psi_h = (sum_q w_q psi_h(invJ (J p_q + v0 - v0))) * measure (integrate_cell, integrator.h:45-59) and
 * D_k = sum of the measures.  Returns the number of triplets (duplicates are left to setFromTriplets). */
int64_t orc_eval_areal(int M, int N, int R, int n_nodes, int n_cells, const double* nodes, const int32_t* cells,
                       const int32_t* dofs, int n_sub, const double* incidence, int32_t* rows, int32_t* cols,
                       double* vals, double* D) {
    orc_fe fe;
    if (fe_init(&fe, M, N, R)) return -1;
    double v[(ORC_MAXD + 1) * ORC_MAXD], J[9], invJ[9], measure, p[ORC_MAXD], xi[ORC_MAXD];
    int64_t tail = 0;
    for (int k = 0; k < n_sub; ++k) {
        int64_t head = 0;
        double Di = 0;
        for (int l = 0; l < n_cells; ++l) {
            if (incidence[(size_t)l * n_sub + k] != 1) continue;
            cell_vertices(M, N, n_nodes, nodes, cells, l, v);
            orc_cell_geometry(M, N, v, J, invJ, &measure);
            for (int h = 0; h < fe.nb; ++h) {
                double value = 0;
                for (int q = 0; q < fe.nq; ++q) {
                    for (int r = 0; r < N; ++r) {
                        double t = 0;
                        for (int m = 0; m < M; ++m) t += J[r * M + m] * fe.qn[q * M + m];
                        p[r] = t + v[r];
                    }
                    for (int m = 0; m < M; ++m) {
                        double t = 0;
                        for (int r = 0; r < N; ++r) t += invJ[m * N + r] * (p[r] - v[r]);
                        xi[m] = t;
                    }
                    value += orc_poly_eval(M, R, fe.coeff + h * fe.nb, xi) * fe.qw[q];
                }
                rows[tail + head] = k;
                cols[tail + head] = dofs[(size_t)h * n_cells + l];
                vals[tail + head] = value * measure;
                ++head;
            }
            Di += measure;
        }
        for (int64_t j = 0; j < head; ++j) vals[tail + j] /= Di;
        D[k] = Di;
        tail += head;
    }
    return tail;
}

/* FEMSolverBase::set_dirichlet_bc (solvers/fem_solver_base.h:144-155) on a CSC matrix: for every boundary dof d
 * (and ALWAYS dof 0: boundary_dofs_begin() returns index 0 untested, :86) zero the stored values of row d,
 * A(d,d) = 1, b(d) = g(d).  The pattern is untouched. */
void orc_set_dirichlet(int n, const int32_t* outer, const int32_t* inner, double* val, const uint8_t* boundary_dofs,
                       const double* g, double* b) {
    for (int j = 0; j < n; ++j)
        for (int32_t k = outer[j]; k < outer[j + 1]; ++k) {
            int i = inner[k];
            if (i == 0 || boundary_dofs[i]) val[k] = (i == j) ? 1.0 : 0.0;
        }
    for (int i = 0; i < n; ++i)
        if (i == 0 || boundary_dofs[i]) b[i] = g[i];
}

/* ------------------------------------------------------------------------------------------
 * Krylov solvers on CSR(A) (NOT in the reference, which only has SparseLU -- SURVEY.md F3; they are the
 * CPU counterpart of the GPU solvers, same stopping rule ||b - A x||_2 / ||b||_2 <= rtol as Eigen's
 * ConjugateGradient/BiCGSTAB use).  Used for the CPU baseline and cross-checks at sizes LU cannot reach.
 * ---------------------------------------------------------------------------------------- */
static void spmv(int n, const int32_t* rp, const int32_t* ci, const double* v, const double* x, double* y) {
    for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int32_t k = rp[i]; k < rp[i + 1]; ++k) s += v[k] * x[ci[k]];
        y[i] = s;
    }
}
static double dot(int n, const double* a, const double* b) {
    double s = 0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

int orc_cg(int n, const int32_t* rp, const int32_t* ci, const double* v, const double* b, double* x, double rtol,
           int maxit, int jacobi, double* rel_resid_out) {
    double *r = malloc(sizeof(double) * n), *p = malloc(sizeof(double) * n), *q = malloc(sizeof(double) * n),
           *z = malloc(sizeof(double) * n), *dinv = malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) {
        dinv[i] = 1.0;
        if (jacobi)
            for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
                if (ci[k] == i && v[k] != 0) dinv[i] = 1.0 / v[k];
    }
    spmv(n, rp, ci, v, x, q);
    for (int i = 0; i < n; ++i) { r[i] = b[i] - q[i]; z[i] = dinv[i] * r[i]; p[i] = z[i]; }
    double bb = dot(n, b, b), rz = dot(n, r, z), rr = dot(n, r, r);
    double thr = rtol * rtol * bb;
    int it = 0;
    if (bb == 0) { for (int i = 0; i < n; ++i) x[i] = 0; rr = 0; }
    while (rr > thr && it < maxit) {
        spmv(n, rp, ci, v, p, q);
        double alpha = rz / dot(n, p, q);
        for (int i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; }
        for (int i = 0; i < n; ++i) z[i] = dinv[i] * r[i];
        double rz_new = dot(n, r, z);
        rr = dot(n, r, r);
        double beta = rz_new / rz;
        rz = rz_new;
        for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
        ++it;
    }
    if (rel_resid_out) *rel_resid_out = bb > 0 ? sqrt(rr / bb) : 0;
    free(r); free(p); free(q); free(z); free(dinv);
    return it;
}

int orc_bicgstab(int n, const int32_t* rp, const int32_t* ci, const double* v, const double* b, double* x,
                 double rtol, int maxit, int jacobi, double* rel_resid_out) {
    double *r = malloc(sizeof(double) * n), *r0 = malloc(sizeof(double) * n), *p = calloc(n, sizeof(double)),
           *vv = calloc(n, sizeof(double)), *s = malloc(sizeof(double) * n), *t = malloc(sizeof(double) * n),
           *y = malloc(sizeof(double) * n), *z = malloc(sizeof(double) * n), *dinv = malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) {
        dinv[i] = 1.0;
        if (jacobi)
            for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
                if (ci[k] == i && v[k] != 0) dinv[i] = 1.0 / v[k];
    }
    spmv(n, rp, ci, v, x, t);
    for (int i = 0; i < n; ++i) { r[i] = b[i] - t[i]; r0[i] = r[i]; }
    double bb = dot(n, b, b), rr = dot(n, r, r), thr = rtol * rtol * bb;
    double rho = 1, alpha = 1, w = 1;
    int it = 0;
    if (bb == 0) { for (int i = 0; i < n; ++i) x[i] = 0; rr = 0; }
    while (rr > thr && it < maxit) {
        double rho_new = dot(n, r0, r);
        if (rho_new == 0) break;
        if (it == 0) { for (int i = 0; i < n; ++i) p[i] = r[i]; }
        else {
            double beta = (rho_new / rho) * (alpha / w);
            for (int i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - w * vv[i]);
        }
        rho = rho_new;
        for (int i = 0; i < n; ++i) y[i] = dinv[i] * p[i];
        spmv(n, rp, ci, v, y, vv);
        alpha = rho / dot(n, r0, vv);
        for (int i = 0; i < n; ++i) s[i] = r[i] - alpha * vv[i];
        for (int i = 0; i < n; ++i) z[i] = dinv[i] * s[i];
        spmv(n, rp, ci, v, z, t);
        double tt = dot(n, t, t);
        w = tt > 0 ? dot(n, t, s) / tt : 0;
        for (int i = 0; i < n; ++i) { x[i] += alpha * y[i] + w * z[i]; r[i] = s[i] - w * t[i]; }
        rr = dot(n, r, r);
        ++it;
    }
    if (rel_resid_out) *rel_resid_out = bb > 0 ? sqrt(rr / bb) : 0;
    free(r); free(r0); free(p); free(vv); free(s); free(t); free(y); free(z); free(dinv);
    return it;
}
