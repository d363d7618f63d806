// textio.cu — native readers of the reference's text formats (host code).
//
// newMultiThreadFit takes a *file name*: the reference reads it line by line and parses every line with
// String.replace / split / toDouble (dense `[id,[v1,v2,...]]`: Vectors.parseDense, src/main/scala/mclab/lsh/vector/
// Vector.scala:215-219, read loop DensevectorRDFInit.scala:172-181; sparse `[id, size, [i1, ...], [v1, ...]]`:
// Vectors.fromPythonString, Vector.scala:194-208, read loop SparsevectorRDFInit.scala:164-176).  With the index build at
// > 100M vectors/s that parse would be the whole fit, so the library reads the file itself: one read of the file, lines
// split across the host's threads, strtod/strtol on the raw bytes, straight into the row-major / CSR arrays the fit
// entry points take.  Semantics kept: blanks are ignored anywhere, the id in the file is ignored (ids are the running
// line number, DensevectorRDFInit.scala:174-184), empty lines are skipped, sparse indices are returned ascending (the
// reference iterates a BitSet, SimilarityCalculator.scala:19-25).
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dpf.h"

namespace {

struct Text {
    std::vector<char> buf;                // whole file + terminating 0
    std::vector<size_t> line_begin;       // non-empty lines only
    std::vector<size_t> line_end;
};

bool blank(char c) { return c == ' ' || c == '\t' || c == '\r'; }

int read_lines(const char* path, Text& t) {
    FILE* f = fopen(path, "rb");
    if (!f) return DPF_ERR_INVALID;
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return DPF_ERR_INVALID; }
    const long sz = ftell(f);
    if (sz < 0) { fclose(f); return DPF_ERR_INVALID; }
    rewind(f);
    t.buf.resize((size_t)sz + 1);
    const size_t got = fread(t.buf.data(), 1, (size_t)sz, f);
    fclose(f);
    if (got != (size_t)sz) return DPF_ERR_INVALID;
    t.buf[(size_t)sz] = 0;
    size_t b = 0;
    for (size_t i = 0; i <= (size_t)sz; ++i) {
        if (i == (size_t)sz || t.buf[i] == '\n') {
            bool any = false;
            for (size_t j = b; j < i && !any; ++j) any = !blank(t.buf[j]);
            if (any) { t.line_begin.push_back(b); t.line_end.push_back(i); }
            if (i < (size_t)sz) t.buf[i] = 0;      // every line is a C string for strtod
            b = i + 1;
        }
    }
    return DPF_OK;
}

template <class F>
void parallel_lines(size_t nlines, F&& body) {
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (nlines < 4096) nt = 1;
    std::vector<std::thread> th;
    for (unsigned w = 0; w < nt; ++w)
        th.emplace_back([&, w] {
            const size_t lo = nlines * w / nt, hi = nlines * (w + 1) / nt;
            for (size_t i = lo; i < hi; ++i) body(i);
        });
    for (auto& x : th) x.join();
}

// next number of the line, skipping blanks, brackets and commas; false at the end of the line
bool next_token(const char*& p) {
    while (*p && (blank(*p) || *p == '[' || *p == ']' || *p == ',' || *p == '(' || *p == ')')) ++p;
    return *p != 0;
}

}  // namespace

extern "C" {

int dpf_parse_dense_file(const char* path, int32_t d, double* X_out, int64_t cap_rows, int64_t* n_out) {
    if (!path || d <= 0 || !n_out) return DPF_ERR_INVALID;
    Text t;
    int rc = read_lines(path, t);
    if (rc != DPF_OK) return rc;
    const int64_t n = (int64_t)t.line_begin.size();
    *n_out = n;
    if (!X_out) return DPF_OK;                         // size query
    if (cap_rows < n) return DPF_ERR_CAPACITY;
    std::atomic<int> bad{0};
    parallel_lines((size_t)n, [&](size_t i) {
        const char* p = t.buf.data() + t.line_begin[i];
        char* e = nullptr;
        if (!next_token(p)) { bad = 1; return; }
        (void)strtol(p, &e, 10);                       // the id in the file is ignored
        if (e == p) { bad = 1; return; }
        p = e;
        double* row = X_out + (int64_t)i * d;
        for (int j = 0; j < d; ++j) {
            if (!next_token(p)) { bad = 1; return; }
            row[j] = strtod(p, &e);
            if (e == p) { bad = 1; return; }
            p = e;
        }
        if (next_token(p)) bad = 1;                    // more than d values
    });
    return bad ? DPF_ERR_INVALID : DPF_OK;
}

int dpf_parse_sparse_file(const char* path, int64_t* indptr_out, int32_t* indices_out, double* values_out, int64_t cap_rows,
                          int64_t cap_nnz, int64_t* n_out, int64_t* nnz_out, int32_t* dim_out) {
    if (!path || !n_out || !nnz_out) return DPF_ERR_INVALID;
    Text t;
    int rc = read_lines(path, t);
    if (rc != DPF_OK) return rc;
    const int64_t n = (int64_t)t.line_begin.size();
    // pass 1: entries per line = numbers inside the first inner bracket pair
    std::vector<int64_t> cnt((size_t)n + 1, 0);
    std::vector<int32_t> dims((size_t)n, 0);
    std::atomic<int> bad{0};
    parallel_lines((size_t)n, [&](size_t i) {
        const char* p = t.buf.data() + t.line_begin[i];
        char* e = nullptr;
        if (!next_token(p)) { bad = 1; return; }
        (void)strtol(p, &e, 10);                       // id (ignored)
        if (e == p) { bad = 1; return; }
        p = e;
        if (!next_token(p)) { bad = 1; return; }
        dims[i] = (int32_t)strtol(p, &e, 10);          // size
        if (e == p) { bad = 1; return; }
        p = e;
        while (*p && *p != '[') ++p;                   // the index list
        if (!*p) { bad = 1; return; }
        ++p;
        int64_t c = 0;
        bool in_num = false;
        for (; *p && *p != ']'; ++p) {
            const bool digit = (*p >= '0' && *p <= '9') || *p == '-' || *p == '+';
            if (digit && !in_num) c++;
            in_num = digit;
        }
        if (!*p) { bad = 1; return; }
        cnt[i + 1] = c;
    });
    if (bad) return DPF_ERR_INVALID;
    for (int64_t i = 0; i < n; ++i) cnt[(size_t)i + 1] += cnt[(size_t)i];
    const int64_t nnz = cnt[(size_t)n];
    *n_out = n;
    *nnz_out = nnz;
    if (dim_out) {
        int32_t mx = 0;
        for (int32_t v : dims) mx = std::max(mx, v);
        *dim_out = mx;
    }
    if (!indptr_out || !indices_out || !values_out) return DPF_OK;      // size query
    if (cap_rows < n || cap_nnz < nnz) return DPF_ERR_CAPACITY;
    for (int64_t i = 0; i <= n; ++i) indptr_out[i] = cnt[(size_t)i];
    parallel_lines((size_t)n, [&](size_t i) {
        const char* p = t.buf.data() + t.line_begin[i];
        char* e = nullptr;
        next_token(p); (void)strtol(p, &e, 10); p = e;
        next_token(p); (void)strtol(p, &e, 10); p = e;
        const int64_t lo = cnt[i], m = cnt[i + 1] - cnt[i];
        std::vector<std::pair<int32_t, double>> ent((size_t)m);
        for (int64_t j = 0; j < m; ++j) {
            if (!next_token(p)) { bad = 1; return; }
            ent[(size_t)j].first = (int32_t)strtol(p, &e, 10);
            if (e == p) { bad = 1; return; }
            p = e;
        }
        for (int64_t j = 0; j < m; ++j) {
            if (!next_token(p)) { bad = 1; return; }
            ent[(size_t)j].second = strtod(p, &e);
            if (e == p) { bad = 1; return; }
            p = e;
        }
        if (next_token(p)) { bad = 1; return; }        // index and value lists of different length
        std::stable_sort(ent.begin(), ent.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        for (int64_t j = 0; j < m; ++j) {
            indices_out[lo + j] = ent[(size_t)j].first;
            values_out[lo + j] = ent[(size_t)j].second;
        }
    });
    return bad ? DPF_ERR_INVALID : DPF_OK;
}

}  // extern "C"
