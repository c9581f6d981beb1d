// Mesh file I/O in the reference's formats, header-only, host side.
//   CSVReader<T>::parse_file<Eigen::Dense>   fdaPDE/utils/IO/csv_reader.h:75-118
//   MeshLoader<Mesh>(meshID)                 test/src/utils/mesh_loader.h:62-84
// One header line; the first column of every row is a row name and is skipped; blanks and '"' are dropped from every
// token; NA / NaN / nan read as NaN.  elements.csv is 1-based on disk.
#ifndef FDAPDE_B200_MESH_IO_H
#define FDAPDE_B200_MESH_IO_H

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "assembler.h"

namespace fdapde_b200 {

// dense CSV -> row-major values + shape
struct CsvTable {
    int rows = 0, cols = 0;
    std::vector<double> values;  // row-major
    double operator()(int i, int j) const { return values[(size_t)i * cols + j]; }
};

inline CsvTable read_csv(const std::string& file) {
    std::ifstream in(file);
    if (!in) throw std::runtime_error("fdapde_b200: cannot open " + file);
    CsvTable t;
    std::string line;
    if (!std::getline(in, line)) return t;
    for (char c : line) t.cols += (c == ',');
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        size_t pos = 0;
        int col = -1;
        while (pos != std::string::npos) {
            size_t next = line.find(',', pos);
            std::string tok;
            for (char c : line.substr(pos, next == std::string::npos ? std::string::npos : next - pos))
                if (c != ' ' && c != '"' && c != '\r') tok += c;
            if (col >= 0 && col < t.cols) {
                if (tok == "NA" || tok == "NaN" || tok == "nan") t.values.push_back(std::numeric_limits<double>::quiet_NaN());
                else t.values.push_back(std::strtod(tok.c_str(), nullptr));
            }
            ++col;
            pos = next == std::string::npos ? next : next + 1;
        }
        for (; col < t.cols; ++col) t.values.push_back(0.0);
        ++t.rows;
    }
    return t;
}

// MeshLoader: <dir>/points.csv, elements.csv (1-based), boundary.csv -> Triangulation<M, N>
template <int M, int N> Triangulation<M, N> load_mesh(const std::string& dir) {
    CsvTable p = read_csv(dir + "/points.csv"), e = read_csv(dir + "/elements.csv"), b = read_csv(dir + "/boundary.csv");
    if (p.cols != N || e.cols != M + 1 || b.rows != p.rows) throw std::runtime_error("fdapde_b200: mesh files do not match <M, N>");
    Triangulation<M, N> m;
    m.n_nodes = p.rows;
    m.n_cells = e.rows;
    m.nodes.resize((size_t)p.rows * N);
    for (int i = 0; i < p.rows; ++i)
        for (int d = 0; d < N; ++d) m.nodes[(size_t)d * p.rows + i] = p(i, d);  // column-major like DMatrix<double>
    m.cells.resize((size_t)e.rows * (M + 1));
    for (int i = 0; i < e.rows; ++i)
        for (int k = 0; k <= M; ++k) m.cells[(size_t)i * (M + 1) + k] = (int32_t)e(i, k) - 1;
    m.boundary.resize(p.rows);
    for (int i = 0; i < p.rows; ++i) m.boundary[i] = b(i, 0) != 0;
    return m;
}

}  // namespace fdapde_b200

#endif  // FDAPDE_B200_MESH_IO_H
