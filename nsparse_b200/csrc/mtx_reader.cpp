// nsparse-b200: MatrixMarket coordinate -> CSR, parallel.
//
// Same result as the reference reader convert_file_csr (cuda-c/src/nsparse.cu:14-136), which parses one
// line at a time with fgets / atoi / atof: first line containing "general" => entries as they are, otherwise
// every off-diagonal entry is mirrored right after itself; '%' lines before the size line are skipped;
// 1-based indices; a missing value reads as 1; entries are appended to their rows in FILE ORDER (nothing is
// sorted or merged); at most `nz` entries (the size line) are read.  Here the file is read once, the body
// is cut at line boundaries into one piece per thread and parsed in parallel (the parse is > 90 % of the
// reference reader's time), and the pieces are concatenated in order before the stable scatter into rows.
// NSP_MTX_SORT_MERGE additionally orders every row by column and sums duplicates (what SpGEMM / AMB want).
#include <limits.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/nsparse_b200.h"

namespace {

struct Piece {
    std::vector<int> r, c;
    std::vector<double> v;
};

inline const char *skip_blank(const char *p)
{
    while (*p == ' ' || *p == '\t') ++p;
    return p;
}

// decimal integer with optional sign; returns false if no digit was read
inline bool parse_long(const char *&p, long &out)
{
    const char *q = skip_blank(p);
    bool neg = false;
    if (*q == '+' || *q == '-') neg = *q++ == '-';
    if (*q < '0' || *q > '9') return false;
    long x = 0;
    while (*q >= '0' && *q <= '9') x = x * 10 + (*q++ - '0');
    out = neg ? -x : x;
    p = q;
    return true;
}

template <typename real>
int build(int M, int N, bool general, int flags, std::vector<Piece> &pieces, long long limit, int *nnz_max,
          long long *nnz_out, int **rpt_out, int **col_out, void **val_out)
{
    // entries in file order, truncated at `limit` originals: eff[t] entries of piece t are used
    const int P = (int)pieces.size();
    std::vector<size_t> eff((size_t)P, 0);
    long long left = limit;
    for (int t = 0; t < P; ++t) {
        eff[t] = (size_t)std::min<long long>((long long)pieces[t].r.size(), left);
        left -= (long long)eff[t];
    }
    std::vector<int> cnt((size_t)M + 1, 0);
    long long nnz = 0;
    int bad = 0;
#pragma omp parallel for schedule(static, 1) reduction(+ : nnz) reduction(| : bad)
    for (int t = 0; t < P; ++t) {
        const Piece &pc = pieces[t];
        for (size_t i = 0; i < eff[t]; ++i) {
            const int r = pc.r[i], c = pc.c[i];
            if (r < 0 || r >= M || c < 0 || c >= N) {
                bad |= 1;
                continue;
            }
#pragma omp atomic
            cnt[r]++;
            ++nnz;
            if (!general && r != c) {
                if (c >= M || r >= N) {
                    bad |= 2;
                    continue;
                }
#pragma omp atomic
                cnt[c]++;
                ++nnz;
            }
        }
    }
    if (bad & 1) {
        fprintf(stderr, "an entry lies outside the %d x %d matrix\n", M, N);
        return NSP_ERR_ARG;
    }
    if (bad & 2) {
        fprintf(stderr, "symmetric file with a non-square shape\n");
        return NSP_ERR_ARG;
    }
    if (nnz > (long long)INT_MAX) {
        fprintf(stderr, "more than 2^31 entries\n");
        return NSP_ERR_OVERFLOW;
    }
    int *rpt = (int *)malloc(sizeof(int) * ((size_t)M + 1));
    int *col = (int *)malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
    real *val = (real *)malloc(sizeof(real) * (size_t)(nnz ? nnz : 1));
    if (!rpt || !col || !val) return NSP_ERR_NOMEM;
    rpt[0] = 0;
    for (int i = 0; i < M; ++i) rpt[i + 1] = rpt[i] + cnt[i];
    // stable scatter, parallel over ROW RANGES of equal size in entries: every thread walks all the entries
    // in file order and places the ones of its rows, so the order inside a row is the file order
    {
        int T = omp_get_max_threads();
        if (nnz < (1 << 16)) T = 1;
        std::vector<int> lo((size_t)T + 1, M);
        lo[0] = 0;
        for (int t = 1; t < T; ++t)
            lo[t] = (int)(std::lower_bound(rpt, rpt + M + 1, (int)(nnz * t / T)) - rpt);
        lo[T] = M;
        for (int t = 1; t <= T; ++t) lo[t] = std::max(lo[t], lo[t - 1]);
#pragma omp parallel for schedule(static, 1) num_threads(T)
        for (int t = 0; t < T; ++t) {
            const int r0 = lo[t], r1 = lo[t + 1];
            if (r0 >= r1) continue;
            std::vector<int> cur(rpt + r0, rpt + r1);
            for (int q = 0; q < P; ++q) {
                const Piece &pc = pieces[q];
                for (size_t i = 0; i < eff[q]; ++i) {
                    const int r = pc.r[i], c = pc.c[i];
                    if (r >= r0 && r < r1) {
                        const int at = cur[r - r0]++;
                        col[at] = c;
                        val[at] = (real)pc.v[i];
                    }
                    if (!general && r != c && c >= r0 && c < r1) {
                        const int at = cur[c - r0]++;
                        col[at] = r;
                        val[at] = (real)pc.v[i];
                    }
                }
            }
        }
    }
    if (flags & NSP_MTX_SORT_MERGE) {
        // per row: stable sort by column, duplicates summed in file order; rows compacted afterwards
        std::vector<int> newlen((size_t)M, 0);
#pragma omp parallel
        {
            std::vector<std::pair<int, real>> tmp;
#pragma omp for schedule(dynamic, 256)
            for (int i = 0; i < M; ++i) {
                const int s = rpt[i], e = rpt[i + 1];
                tmp.clear();
                for (int k = s; k < e; ++k) tmp.emplace_back(col[k], val[k]);
                std::stable_sort(tmp.begin(), tmp.end(), [](const std::pair<int, real> &a, const std::pair<int, real> &b) {
                    return a.first < b.first;
                });
                int w = s;
                for (size_t k = 0; k < tmp.size(); ++k) {
                    if (w > s && col[w - 1] == tmp[k].first)
                        val[w - 1] += tmp[k].second;
                    else {
                        col[w] = tmp[k].first;
                        val[w] = tmp[k].second;
                        ++w;
                    }
                }
                newlen[i] = w - s;
            }
        }
        int w = 0;
        for (int i = 0; i < M; ++i) {
            const int s = rpt[i];
            rpt[i] = w;
            if (w != s) {
                memmove(col + w, col + s, sizeof(int) * (size_t)newlen[i]);
                memmove(val + w, val + s, sizeof(real) * (size_t)newlen[i]);
            }
            w += newlen[i];
        }
        rpt[M] = w;
        nnz = w;
    }
    int mx = 0;
    for (int i = 0; i < M; ++i) mx = std::max(mx, rpt[i + 1] - rpt[i]);
    *nnz_max = mx;
    *nnz_out = nnz;
    *rpt_out = rpt;
    *col_out = col;
    *val_out = val;
    return 0;
}

}  // namespace

extern "C" {

int nsp_read_mtx(const char *path, int is_double, int flags, int *M_out, int *N_out, long long *nnz_out,
                 int *nnz_max_out, int **rpt_out, int **col_out, void **val_out)
{
    if (!path || !M_out || !N_out || !nnz_out || !nnz_max_out || !rpt_out || !col_out || !val_out) return NSP_ERR_ARG;
    FILE *fp = fopen(path, "rb");
    if (!fp) return NSP_ERR_ARG;
    fseek(fp, 0, SEEK_END);
    const long fsize = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<char> buf((size_t)fsize + 2);
    if (fsize > 0 && fread(buf.data(), 1, (size_t)fsize, fp) != (size_t)fsize) {
        fclose(fp);
        return NSP_ERR_ARG;
    }
    fclose(fp);
    buf[fsize] = '\n';
    buf[fsize + 1] = 0;
    const char *p = buf.data(), *end = buf.data() + fsize;
    auto next_line = [&](const char *q) {
        const char *nl = (const char *)memchr(q, '\n', (size_t)(end + 1 - q));
        return nl ? nl + 1 : end + 1;
    };
    // banner: "general" anywhere in the first line
    const char *l1 = next_line(p);
    const bool general = std::search(p, l1, "general", "general" + 7) != l1;
    p = l1;
    while (p <= end && *p == '%') p = next_line(p);
    if (p > end) return NSP_ERR_ARG;
    int M = 0, N = 0, nz = 0;
    if (sscanf(p, "%d %d %d", &M, &N, &nz) != 3 || M < 0 || N < 0 || nz < 0) return NSP_ERR_ARG;
    p = next_line(p);
    const char *body = p <= end ? p : end + 1;
    const long blen = (long)(end + 1 - body);
    int T = omp_get_max_threads();
    if (blen < (1 << 16)) T = 1;
    std::vector<const char *> cut((size_t)T + 1);
    cut[0] = body;
    for (int t = 1; t < T; ++t) {
        const char *q = body + blen * t / T;
        cut[t] = q <= body ? body : next_line(q - 1);       // first line start at or after q
        if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    }
    cut[T] = end + 1;
    std::vector<Piece> pieces((size_t)T);
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t) {
        Piece &pc = pieces[t];
        const char *q = cut[t];
        const char *stop = cut[t + 1];
        const size_t guess = (size_t)(stop - q) / 12 + 16;
        pc.r.reserve(guess);
        pc.c.reserve(guess);
        pc.v.reserve(guess);
        while (q < stop) {
            const char *nl = next_line(q);
            const char *s = q;
            long r, c;
            if (parse_long(s, r) && parse_long(s, c)) {
                // a third word on this line is the value (atof: 0 if it is not a number), none means 1
                const char *w = skip_blank(s);
                double v = 1.0;
                if (*w != '\n' && *w != '\r' && *w != 0) v = strtod(w, nullptr);
                pc.r.push_back((int)(r - 1));
                pc.c.push_back((int)(c - 1));
                pc.v.push_back(v);
            }
            q = nl;
        }
    }
    *M_out = M;
    *N_out = N;
    if (is_double)
        return build<double>(M, N, general, flags, pieces, nz, nnz_max_out, nnz_out, rpt_out, col_out, val_out);
    return build<float>(M, N, general, flags, pieces, nz, nnz_max_out, nnz_out, rpt_out, col_out, val_out);
}

void nsp_free_host(void *p) { free(p); }

}  // extern "C"
