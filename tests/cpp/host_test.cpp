// Host-only part of the C++ shim (no device call): Eigen's setFromTriplets semantics in sp_from_triplets and the
// reference's CSV / MeshLoader formats in mesh_io.h.  Runs on the CPU box.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "fdapde_b200/assembler.h"
#include "fdapde_b200/mesh_io.h"

using namespace fdapde_b200;

static int failures = 0;
#define EXPECT_TRUE(c)                                                 \
    do {                                                               \
        if (!(c)) {                                                    \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            ++failures;                                                \
        }                                                              \
    } while (0)

// setFromTriplets (fem_assembler.h:112-113): column-major, inner indices sorted, duplicates summed in list order,
// explicit zeros kept
static void triplets() {
    std::vector<int32_t> r = {2, 0, 2, 1, 2, 0}, c = {1, 0, 1, 2, 1, 0};
    std::vector<double> v = {1e16, 3.0, 1.0, 0.0, -1e16, 4.0};
    SpMatrix A = sp_from_triplets(3, 3, r, c, v);
    EXPECT_TRUE(A.rows == 3 && A.cols == 3 && A.nonZeros() == 3);
    EXPECT_TRUE((A.outer == std::vector<int32_t>{0, 1, 2, 3}));
    EXPECT_TRUE((A.inner == std::vector<int32_t>{0, 2, 1}));
    EXPECT_TRUE(A.coeff(0, 0) == 7.0);
    EXPECT_TRUE(A.coeff(2, 1) == (1e16 + 1.0) + -1e16);   // left to right: the 1.0 is absorbed, result 0
    EXPECT_TRUE(A.values[2] == 0.0 && A.inner[2] == 1);    // the explicit zero (1, 2) is stored
    // unsorted rows inside a column come out sorted
    SpMatrix B = sp_from_triplets(4, 1, {3, 1, 2, 1}, {0, 0, 0, 0}, {1, 2, 3, 4});
    EXPECT_TRUE((B.inner == std::vector<int32_t>{1, 2, 3}) && B.values[0] == 6.0);
    SpMatrix E = sp_from_triplets(2, 2, {}, {}, {});
    EXPECT_TRUE(E.nonZeros() == 0 && E.outer.size() == 3);
}

static void csv_and_mesh() {
    const char* tmp = std::getenv("TMPDIR");
    const std::string dir = std::string(tmp ? tmp : "/tmp") + "/fdb_host_test_mesh";
    if (std::system(("mkdir -p " + dir).c_str()) != 0) { ++failures; return; }
    {
        std::ofstream p(dir + "/points.csv"), e(dir + "/elements.csv"), b(dir + "/boundary.csv");
        p << "\"\",\"V1\",\"V2\"\n\"1\",\"0\",\" 0\"\n\"2\",\"1\",\"0\"\n\"3\",\"0\",\"1\"\n\"4\",\"1.0000000000000002\",NA\n";
        e << "\"\",\"V1\",\"V2\",\"V3\"\n\"1\",1,2,3\n\"2\",2,4,3\n";
        b << "\"\",\"V1\"\n\"1\",1\n\"2\",1\n\"3\",1\n\"4\",0\n";
    }
    CsvTable t = read_csv(dir + "/points.csv");
    EXPECT_TRUE(t.rows == 4 && t.cols == 2);
    EXPECT_TRUE(t(1, 0) == 1.0 && t(3, 0) == 1.0000000000000002 && std::isnan(t(3, 1)));
    auto m = load_mesh<2, 2>(dir);
    EXPECT_TRUE(m.n_nodes == 4 && m.n_cells == 2);
    EXPECT_TRUE((m.cells == std::vector<int32_t>{0, 1, 2, 1, 3, 2}));   // 0-based in memory
    EXPECT_TRUE(m.nodes[1] == 1.0 && m.nodes[4 + 2] == 1.0);           // column-major: x then y
    EXPECT_TRUE(m.boundary[0] == 1 && m.boundary[3] == 0);
    bool threw = false;
    try { load_mesh<3, 3>(dir); } catch (const std::runtime_error&) { threw = true; }
    EXPECT_TRUE(threw);
}

// the operator expression algebra lowers like differential_expressions.h:54-135
static void expressions() {
    auto L = -laplacian<FEM>() + 2.0 * reaction<FEM>(3.0) - advection<FEM>(std::vector<double>{1.0, -1.0});
    EXPECT_TRUE(!L.is_symmetric());
    EXPECT_TRUE((-laplacian<FEM>() + reaction<FEM>(1.0)).is_symmetric());
    fdb_opdesc d;
    L.lower(&d);
    EXPECT_TRUE(d.n_terms == 3 && d.symmetric == 0);
    EXPECT_TRUE(d.terms[0].kind == FDB_LAPLACIAN && d.terms[0].scale == -1.0);
    EXPECT_TRUE(d.terms[1].kind == FDB_REACTION && d.terms[1].scale == 2.0);
    EXPECT_TRUE(d.terms[2].kind == FDB_ADVECTION && d.terms[2].scale == -1.0);
}

int main() {
    triplets();
    csv_and_mesh();
    expressions();
    std::printf(failures ? "HOST_TEST_FAIL\n" : "HOST_TEST_PASS\n");
    return failures ? 1 : 0;
}
