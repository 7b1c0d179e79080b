// MatrixMarket.cpp -- Matrix Market coordinate files <-> the COO / CSR field layout of the matrix classes.
//
// The reference names this as the first of its planned utilities ("reading in matrices from file formats like MATLAB
// or Matrix Market and storing them in standard formats (e.g. COO, CSR, ELL, ...)", README.md:90-99) and ships no
// reader; its test programs only generate stencils.  Host-only code (no CUDA): the arrays it fills are what
// lsk_coo_create / lsk_csr_create upload.
//
// Supported: `%%MatrixMarket matrix coordinate {real|integer|pattern} {general|symmetric|skew-symmetric}`.
// Indices become 0-based; symmetric / skew-symmetric storage is expanded (the mirrored entry follows its original);
// pattern entries get the value 1.  `array` (dense) and `complex` / `hermitian` files are refused.
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/lsk_solvers.h"

namespace {

struct Header {
    lsk_mm_info info{};
    std::streampos body;  // first entry line
};

std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char) std::tolower(c); });
    return s;
}

Header read_header(std::ifstream &in, const std::string &path) {
    if (!in) throw std::runtime_error("cannot open " + path);
    std::string line;
    if (!std::getline(in, line)) throw std::runtime_error(path + ": empty file");
    std::istringstream banner(line);
    std::string tag, object, format, field, symmetry;
    banner >> tag >> object >> format >> field >> symmetry;
    if (tag != "%%MatrixMarket") throw std::runtime_error(path + ": not a Matrix Market file (no %%MatrixMarket banner)");
    object = lower(object); format = lower(format); field = lower(field); symmetry = lower(symmetry);
    if (object != "matrix") throw std::runtime_error(path + ": object '" + object + "' is not supported (matrix only)");
    if (format != "coordinate") throw std::runtime_error(path + ": format '" + format + "' is not supported (coordinate only)");
    Header h;
    if (field == "real") h.info.field = LSK_MM_REAL;
    else if (field == "integer") h.info.field = LSK_MM_INTEGER;
    else if (field == "pattern") h.info.field = LSK_MM_PATTERN;
    else throw std::runtime_error(path + ": field '" + field + "' is not supported (real, integer, pattern)");
    if (symmetry == "general") h.info.symmetry = LSK_MM_GENERAL;
    else if (symmetry == "symmetric") h.info.symmetry = LSK_MM_SYMMETRIC;
    else if (symmetry == "skew-symmetric") h.info.symmetry = LSK_MM_SKEW_SYMMETRIC;
    else throw std::runtime_error(path + ": symmetry '" + symmetry + "' is not supported (general, symmetric, skew-symmetric)");
    // comments and blank lines, then the size line
    for (;;) {
        if (!std::getline(in, line)) throw std::runtime_error(path + ": no size line");
        size_t i = 0;
        while (i < line.size() && std::isspace((unsigned char) line[i])) ++i;
        if (i == line.size() || line[i] == '%') continue;
        break;
    }
    long long r = -1, c = -1, e = -1;
    if (std::sscanf(line.c_str(), "%lld %lld %lld", &r, &c, &e) != 3 || r < 0 || c < 0 || e < 0)
        throw std::runtime_error(path + ": malformed size line '" + line + "'");
    if (h.info.symmetry != LSK_MM_GENERAL && r != c) throw std::runtime_error(path + ": symmetric storage of a non-square matrix");
    h.info.rows = r;
    h.info.cols = c;
    h.info.entries = e;
    h.body = in.tellg();
    return h;
}

}  // namespace

static thread_local std::string g_mm_error;

template <class F>
static int mm_guard(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_mm_error = e.what();
        return LSK_E_INVALID;
    }
}

extern "C" {

const char *lsk_mm_last_error(void) { return g_mm_error.c_str(); }

int lsk_mm_read_info(const char *path, lsk_mm_info *out) {
    if (!path || !out) return LSK_E_INVALID;
    return mm_guard([&] {
        std::ifstream in(path);
        *out = read_header(in, path).info;
    });
}

int lsk_mm_read_coo_f64(const char *path, int64_t capacity, double *entry, int64_t *row, int64_t *col, int64_t *nnz_out) {
    if (!path || !nnz_out || capacity < 0 || (capacity > 0 && (!entry || !row || !col))) return LSK_E_INVALID;
    *nnz_out = 0;
    return mm_guard([&] {
        std::ifstream in(path);
        const Header h = read_header(in, path);
        const lsk_mm_info &mi = h.info;
        int64_t n = 0;
        auto put = [&](int64_t r, int64_t c, double v) {
            if (n >= capacity) throw std::runtime_error(std::string(path) + ": more entries than the arrays hold (capacity " + std::to_string(capacity) + ")");
            row[n] = r; col[n] = c; entry[n] = v;
            ++n;
        };
        std::string line;
        int64_t seen = 0;
        while (seen < mi.entries) {
            if (!std::getline(in, line)) throw std::runtime_error(std::string(path) + ": " + std::to_string(seen) + " entries instead of the announced " + std::to_string(mi.entries));
            const char *p = line.c_str();
            while (*p && std::isspace((unsigned char) *p)) ++p;
            if (!*p || *p == '%') continue;
            char *end = nullptr;
            const long long r = std::strtoll(p, &end, 10);
            if (end == p) throw std::runtime_error(std::string(path) + ": malformed entry '" + line + "'");
            p = end;
            const long long c = std::strtoll(p, &end, 10);
            if (end == p) throw std::runtime_error(std::string(path) + ": malformed entry '" + line + "'");
            p = end;
            double v = 1.0;
            if (mi.field != LSK_MM_PATTERN) {
                v = std::strtod(p, &end);
                if (end == p) throw std::runtime_error(std::string(path) + ": entry without a value '" + line + "'");
            }
            if (r < 1 || r > mi.rows || c < 1 || c > mi.cols) throw std::runtime_error(std::string(path) + ": index out of range in '" + line + "'");
            put(r - 1, c - 1, v);
            if (mi.symmetry != LSK_MM_GENERAL && r != c) put(c - 1, r - 1, mi.symmetry == LSK_MM_SKEW_SYMMETRIC ? -v : v);
            if (mi.symmetry == LSK_MM_SKEW_SYMMETRIC && r == c) throw std::runtime_error(std::string(path) + ": diagonal entry in a skew-symmetric file");
            ++seen;
        }
        *nnz_out = n;
    });
}

int lsk_mm_write_coo_f64(const char *path, int64_t rows, int64_t cols, int64_t nnz, const double *entry, const int64_t *row, const int64_t *col) {
    if (!path || rows < 0 || cols < 0 || nnz < 0 || (nnz > 0 && (!entry || !row || !col))) return LSK_E_INVALID;
    return mm_guard([&] {
        FILE *f = std::fopen(path, "w");
        if (!f) throw std::runtime_error(std::string("cannot create ") + path);
        std::fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n%% written by legionsolvers_b200\n%lld %lld %lld\n", (long long) rows, (long long) cols,
                     (long long) nnz);
        bool ok = true;
        for (int64_t k = 0; k < nnz && ok; ++k) {
            if (row[k] < 0 || row[k] >= rows || col[k] < 0 || col[k] >= cols) ok = false;
            else std::fprintf(f, "%lld %lld %.17g\n", (long long) row[k] + 1, (long long) col[k] + 1, entry[k]);  // 17 digits: doubles round-trip
        }
        const bool closed = std::fclose(f) == 0;
        if (!ok) throw std::runtime_error(std::string(path) + ": index out of range");
        if (!closed) throw std::runtime_error(std::string("error writing ") + path);
    });
}

// COO (any order) -> the reference's CSR field layout: entries grouped by row (a stable counting sort: the order of a row's
// entries is their order in the input), col, and one INCLUSIVE rect {first k, last k} per row (empty row: lo > hi).
int lsk_coo_to_csr_f64(int64_t rows, int64_t nnz, const double *entry, const int64_t *row, const int64_t *col, double *entry_out, int64_t *col_out,
                       lsk_rect *rowptr_out) {
    if (rows < 0 || nnz < 0 || (rows > 0 && !rowptr_out) || (nnz > 0 && (!entry || !row || !col || !entry_out || !col_out))) return LSK_E_INVALID;
    return mm_guard([&] {
        std::vector<int64_t> start((size_t) rows + 1, 0);
        for (int64_t k = 0; k < nnz; ++k) {
            if (row[k] < 0 || row[k] >= rows) throw std::runtime_error("lsk_coo_to_csr_f64: row index out of range");
            start[(size_t) row[k] + 1] += 1;
        }
        for (int64_t r = 0; r < rows; ++r) start[(size_t) r + 1] += start[(size_t) r];
        for (int64_t r = 0; r < rows; ++r) {
            rowptr_out[r].lo = start[(size_t) r];
            rowptr_out[r].hi = start[(size_t) r + 1] - 1;
        }
        std::vector<int64_t> next(start.begin(), start.end() - 1);
        for (int64_t k = 0; k < nnz; ++k) {
            const int64_t d = next[(size_t) row[k]]++;
            entry_out[d] = entry[k];
            col_out[d] = col[k];
        }
    });
}

}  // extern "C"
