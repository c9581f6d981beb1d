// Reference-element data of the Lagrange P1/P2 spaces, evaluated once on the host.
//
// Replaces, for the GPU path, what the reference re-evaluates per (cell, i, j, quadrature node):
//   LagrangianElement ctor            basis/lagrangian_basis.h:50-91   (Vandermonde system on the reference nodes)
//   MultivariatePolynomial / derive   basis/multivariate_polynomial.h:157-216
//   ReferenceElement<M,R>::nodes      basis/reference_element.h:28-97
//   IntegratorTable<M,K>              utils/integration/integrator_tables.h:64-320 (rule choice :23-58)
// The quadrature constants are the reference's tabulated 15-digit values on purpose: results must match the
// reference's, not the textbook rule.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace fdb {

static int n_monomials(int M, int R) {
    int num = 1, den = 1;
    for (int k = 1; k <= R; ++k) {
        num *= (M + k);
        den *= k;
    }
    return num / den;
}

// exponents of total degree <= R, first coordinate running fastest (multivariate_polynomial.h:52-79)
static void exponent_table(int M, int R, int* e) {
    int nm = 0;
    int hi1 = (M >= 2) ? R : 0, hi2 = (M >= 3) ? R : 0;
    for (int c = 0; c <= hi2; ++c)
        for (int b = 0; b <= hi1; ++b)
            for (int a = 0; a <= R; ++a)
                if (a + b + c <= R) {
                    int ex[3] = {a, b, c};
                    for (int k = 0; k < M; ++k) e[nm * M + k] = ex[k];
                    ++nm;
                }
}

static double ipow(double x, int e) {
    double r = 1.0;
    for (int k = 0; k < e; ++k) r *= x;
    return r;
}
static double monomial(int M, const double* p, const int* e) {
    double m = 1.0;
    for (int k = 0; k < M; ++k) m *= ipow(p[k], e[k]);
    return m;
}

static int reference_nodes(int M, int R, double* out) {
    // vertices first, then (R == 2) the edge midpoints in the order of ReferenceElement<M,2>::nodes
    static const double t22[] = {0, 0, 1, 0, 0, 1, 0.5, 0, 0, 0.5, 0.5, 0.5};
    static const double t32[] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0.5, 0.5, 0, 0, 0.5, 0, 0.5, 0, 0,
                                 0.5, 0, 0.5, 0, 0.5, 0.5, 0, 0, 0.5};
    if (R == 1) {
        for (int i = 0; i <= M; ++i)
            for (int k = 0; k < M; ++k) out[i * M + k] = (i == k + 1) ? 1.0 : 0.0;
        return 0;
    }
    if (R == 2 && M == 2) { memcpy(out, t22, sizeof(t22)); return 0; }
    if (R == 2 && M == 3) { memcpy(out, t32, sizeof(t32)); return 0; }
    return -1;
}

static int quadrature_rule(int M, int R, int* nq, double* nodes, double* w) {
    static const double n23[] = {0.166666666666667, 0.166666666666667, 0.666666666666667,
                                 0.166666666666667, 0.166666666666667, 0.666666666666667};
    static const double w23[] = {0.333333333333333, 0.333333333333333, 0.333333333333333};
    static const double n26[] = {0.445948490915965, 0.445948490915965, 0.445948490915965, 0.108103018168070,
                                 0.108103018168070, 0.445948490915965, 0.091576213509771, 0.091576213509771,
                                 0.091576213509771, 0.816847572980459, 0.816847572980459, 0.091576213509771};
    static const double w26[] = {0.223381589678011, 0.223381589678011, 0.223381589678011,
                                 0.109951743655322, 0.109951743655322, 0.109951743655322};
    static const double n34[] = {0.585410196624969, 0.138196601125011, 0.138196601125011, 0.138196601125011,
                                 0.138196601125011, 0.138196601125011, 0.138196601125011, 0.138196601125011,
                                 0.585410196624969, 0.138196601125011, 0.585410196624969, 0.138196601125011};
    static const double w34[] = {0.250000000000000, 0.250000000000000, 0.250000000000000, 0.250000000000000};
    static const double n35[] = {0.250000000000000, 0.250000000000000, 0.250000000000000, 0.500000000000000,
                                 0.166666666666667, 0.166666666666667, 0.166666666666667, 0.500000000000000,
                                 0.166666666666667, 0.166666666666667, 0.166666666666667, 0.500000000000000,
                                 0.166666666666667, 0.166666666666667, 0.166666666666667};
    static const double w35[] = {-0.80000000000000, 0.450000000000000, 0.450000000000000, 0.450000000000000,
                                 0.450000000000000};
    const double *n = nullptr, *ww = nullptr;
    if (M == 2 && R == 1) { *nq = 3; n = n23; ww = w23; }
    if (M == 2 && R == 2) { *nq = 6; n = n26; ww = w26; }
    if (M == 3 && R == 1) { *nq = 4; n = n34; ww = w34; }
    if (M == 3 && R == 2) { *nq = 5; n = n35; ww = w35; }
    if (!n) return -1;
    memcpy(nodes, n, sizeof(double) * (*nq) * M);
    memcpy(w, ww, sizeof(double) * (*nq));
    return 0;
}

int build_fe_tables(int M, int R, FeTables* t, PolyTables* poly) {
    memset(t, 0, sizeof(*t));
    FDB_CHECK((M == 2 || M == 3) && (R == 1 || R == 2), FDB_ERR_UNSUPPORTED,
              "only M in {2,3} and R in {1,2} are supported");
    t->M = M;
    t->R = R;
    int nb = n_monomials(M, R);
    t->nb = nb;
    FDB_CHECK(quadrature_rule(M, R, &t->nq, t->qn, t->w) == 0, FDB_ERR_UNSUPPORTED, "no quadrature rule");
    FDB_CHECK(reference_nodes(M, R, t->refn) == 0, FDB_ERR_UNSUPPORTED, "no reference element");
    int ex[MAX_NB * MAX_D];
    exponent_table(M, R, ex);
    // Vandermonde V[i][m] = node_i^{e_m}; basis coefficients = columns of V^{-1} (Gauss-Jordan, row pivoting)
    double a[MAX_NB][2 * MAX_NB];
    for (int i = 0; i < nb; ++i)
        for (int m = 0; m < nb; ++m) {
            a[i][m] = monomial(M, t->refn + i * M, ex + m * M);
            a[i][nb + m] = (i == m) ? 1.0 : 0.0;
        }
    for (int k = 0; k < nb; ++k) {
        int piv = k;
        for (int r = k + 1; r < nb; ++r)
            if (std::fabs(a[r][k]) > std::fabs(a[piv][k])) piv = r;
        FDB_CHECK(a[piv][k] != 0.0, FDB_ERR_ARG, "singular Vandermonde matrix");
        if (piv != k)
            for (int c = 0; c < 2 * nb; ++c) std::swap(a[k][c], a[piv][c]);
        double d = a[k][k];
        for (int c = 0; c < 2 * nb; ++c) a[k][c] /= d;
        for (int r = 0; r < nb; ++r)
            if (r != k && a[r][k] != 0.0) {
                double f = a[r][k];
                for (int c = 0; c < 2 * nb; ++c) a[r][c] -= f * a[k][c];
            }
    }
    // coefficient of monomial m in psi_i = Vinv[m][i]
    if (poly) {
        memset(poly, 0, sizeof(*poly));
        poly->M = M; poly->R = R; poly->nb = nb;
        for (int m = 0; m < nb; ++m) {
            for (int d = 0; d < M; ++d) poly->ex[m * M + d] = ex[m * M + d];
            for (int i = 0; i < nb; ++i) poly->coef[i * nb + m] = a[m][nb + i];
        }
    }
    for (int q = 0; q < t->nq; ++q) {
        const double* p = t->qn + q * M;
        for (int i = 0; i < nb; ++i) {
            double v = 0;
            double g[MAX_D] = {0, 0, 0};
            for (int m = 0; m < nb; ++m) {
                double c = a[m][nb + i];
                v += c * monomial(M, p, ex + m * M);
                for (int d = 0; d < M; ++d) {
                    int e = ex[m * M + d];
                    if (e == 0) continue;
                    int ge[MAX_D];
                    for (int z = 0; z < M; ++z) ge[z] = ex[m * M + z] - (z == d ? 1 : 0);
                    g[d] += c * e * monomial(M, p, ge);
                }
            }
            t->phi[q * nb + i] = v;
            for (int d = 0; d < M; ++d) t->gref[(q * nb + i) * M + d] = g[d];
        }
    }
    return FDB_OK;
}

}  // namespace fdb
