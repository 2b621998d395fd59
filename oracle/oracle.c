/*
 * oracle.c -- CPU restatement of the nsparse hot paths.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (nsparse_b200/) never links, imports or calls anything in oracle/.
 *
 * What it restates (paths relative to the nsparse reference tree):
 *   orc_mtx_read          cuda-c/src/nsparse.cu:14-136   convert_file_csr (MatrixMarket -> CSR)
 *   orc_spgemm_flop       cuda-c/src/kernel/kernel_spgemm_cu_csr.cu:18-33,52-54
 *   orc_spgemm_symbolic   cuda-c/src/kernel/kernel_spgemm_hash_d.cu:399-472 (+1183 scan)
 *   orc_spgemm_numeric_*  cuda-c/src/kernel/kernel_spgemm_hash_d.cu:829-927
 *   orc_spmv_csr_*        cuda-c/src/nsparse.cu:240-259   csr_kernel
 *
 * The reference has NO CPU SpGEMM and its own SpGEMM oracle (legacy cusparse?csrgemm,
 * kernel_spgemm_cu_csr.cu:99,134,142) was removed from cuSPARSE, so for SpGEMM this file is the
 * oracle of record: row-wise Gustavson with the reference's hash function ((col*107) & (size-1),
 * linear probing), structural nnz (numerical zeros kept), rows sorted by ascending column.
 * Pinned by: the golden vectors of data/test.mtx (tests/golden/), SciPy (tests), and
 * for the reader and the CPU SpMV the reference's own nsparse.cu compiled into oracle/_ref/.
 *
 * Parallelism: OpenMP over rows, one private table per thread; per-row arithmetic is sequential in
 * the order the reference's single thread would see the products, so results do not depend on the
 * thread count.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_LINE_MAX 256 /* LINE_LENGTH_MAX, nsparse.cu:11 */
#define ORC_HASH_SCAL 107u /* HASH_SCAL, kernel_spgemm_hash_d.cu:30 */

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * MatrixMarket reader, nsparse.cu:14-136.  Same observable behaviour:
 *   - first line containing "general" => entries are taken as they are, otherwise every
 *     off-diagonal entry is mirrored (:40-42, :87-90, :118-121);
 *   - lines starting with '%' are skipped, the first other line is "M N nz" (:43-48);
 *   - row/col are 1-based atoi() of the first two space-separated words; the value is atof() of
 *     the third word, 1.0 when there is none (:58-76);
 *   - entries are appended to their rows in FILE ORDER, mirrored entries right after the
 *     original; nothing is sorted or merged (:114-123);
 *   - nnz_max = longest row (:103-105).
 * Deviation: a line that does not start with a digit (e.g. a trailing blank line) is ignored
 * instead of being parsed as entry (0,0); *nz is checked against the buffer.
 * is_double selects float/double storage of val.  Returns 0 on success.
 * ------------------------------------------------------------------------------------------ */
int orc_mtx_read(const char *path, int is_double, int *M, int *N, int *nnz, int *nnz_max, int **rpt_out,
                 int **col_out, void **val_out)
{
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    char line[ORC_LINE_MAX];
    int is_unsym = 0;
    if (!fgets(line, ORC_LINE_MAX, fp)) {
        fclose(fp);
        return -2;
    }
    if (strstr(line, "general")) is_unsym = 1;
    do {
        if (!fgets(line, ORC_LINE_MAX, fp)) {
            fclose(fp);
            return -2;
        }
    } while (line[0] == '%');
    int nz = 0;
    if (sscanf(line, "%d %d %d", M, N, &nz) != 3) {
        fclose(fp);
        return -2;
    }
    int *row_coo = (int *)malloc(sizeof(int) * (size_t)(nz > 0 ? nz : 1));
    int *col_coo = (int *)malloc(sizeof(int) * (size_t)(nz > 0 ? nz : 1));
    double *val_coo = (double *)malloc(sizeof(double) * (size_t)(nz > 0 ? nz : 1));
    int num = 0;
    while (fgets(line, ORC_LINE_MAX, fp) && num < nz) {
        char *ch = line;
        if (!(ch[0] >= '0' && ch[0] <= '9')) continue;
        row_coo[num] = atoi(ch) - 1;
        ch = strchr(ch, ' ');
        if (!ch) continue;
        ch++;
        col_coo[num] = atoi(ch) - 1;
        ch = strchr(ch, ' ');
        if (ch != NULL) {
            ch++;
            /* the reference casts atof() to `real` here (:68); done at the store below */
            val_coo[num] = atof(ch);
        } else {
            val_coo[num] = 1.0;
        }
        num++;
    }
    fclose(fp);

    int total = num;
    int *cnt = (int *)calloc((size_t)(*M > 0 ? *M : 1), sizeof(int));
    for (int i = 0; i < num; i++) {
        cnt[row_coo[i]]++;
        if (col_coo[i] != row_coo[i] && !is_unsym) {
            cnt[col_coo[i]]++;
            total++;
        }
    }
    int *rpt = (int *)malloc(sizeof(int) * ((size_t)*M + 1));
    int *col = (int *)malloc(sizeof(int) * (size_t)(total > 0 ? total : 1));
    void *val = malloc((is_double ? sizeof(double) : sizeof(float)) * (size_t)(total > 0 ? total : 1));
    int off = 0, mx = 0;
    for (int i = 0; i < *M; i++) {
        rpt[i] = off;
        off += cnt[i];
        if (cnt[i] > mx) mx = cnt[i];
    }
    rpt[*M] = off;
    int *fill = (int *)calloc((size_t)(*M > 0 ? *M : 1), sizeof(int));
    for (int i = 0; i < num; i++) {
        int r = row_coo[i], c = col_coo[i];
        int p = rpt[r] + fill[r]++;
        col[p] = c;
        if (is_double) ((double *)val)[p] = val_coo[i];
        else ((float *)val)[p] = (float)val_coo[i];
        if (c != r && !is_unsym) {
            p = rpt[c] + fill[c]++;
            col[p] = r;
            if (is_double) ((double *)val)[p] = val_coo[i];
            else ((float *)val)[p] = (float)val_coo[i];
        }
    }
    free(row_coo);
    free(col_coo);
    free(val_coo);
    free(cnt);
    free(fill);
    *nnz = total;
    *nnz_max = mx;
    *rpt_out = rpt;
    *col_out = col;
    *val_out = val;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * flop count, kernel_spgemm_cu_csr.cu:27-32,54:  2 * sum_i sum_{j in A_i} nnz(B_j)
 * ------------------------------------------------------------------------------------------ */
long long orc_spgemm_flop(int M, const int *a_rpt, const int *a_col, const int *b_rpt)
{
    long long total = 0;
#pragma omp parallel for reduction(+ : total) schedule(dynamic, 1024)
    for (int i = 0; i < M; i++) {
        long long s = 0;
        for (int j = a_rpt[i]; j < a_rpt[i + 1]; j++) s += b_rpt[a_col[j] + 1] - b_rpt[a_col[j]];
        total += s;
    }
    return 2 * total;
}

/* per-row intermediate products (set_intprod_num, kernel_spgemm_hash_d.cu:70-86) */
void orc_spgemm_intprod(int M, const int *a_rpt, const int *a_col, const int *b_rpt, long long *ip)
{
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < M; i++) {
        long long s = 0;
        for (int j = a_rpt[i]; j < a_rpt[i + 1]; j++) s += b_rpt[a_col[j] + 1] - b_rpt[a_col[j]];
        ip[i] = s;
    }
}

static size_t pow2_at_least(long long v)
{
    size_t p = 32;
    while ((long long)p < v) p <<= 1;
    return p;
}

/* ------------------------------------------------------------------------------------------
 * SYMBOLIC: nnz of each row of C by inserting every product's column into a per-row hash set
 * (probe loop of kernel_spgemm_hash_d.cu:426-448), then the exclusive scan of :1183.
 * The table holds 2x the row's intermediate products (the reference sizes it 1x..8x by bin);
 * the count does not depend on the table size.  rows [row_begin,row_end) only; c_rpt has
 * row_end-row_begin+1 entries and starts at 0.
 * ------------------------------------------------------------------------------------------ */
long long orc_spgemm_symbolic_n(int row_begin, int row_end, const int *a_rpt, const int *a_col, const int *b_rpt,
                                const int *b_col, long long *c_rpt, int n_cols);

long long orc_spgemm_symbolic(int row_begin, int row_end, const int *a_rpt, const int *a_col, const int *b_rpt,
                              const int *b_col, long long *c_rpt)
{
    return orc_spgemm_symbolic_n(row_begin, row_end, a_rpt, a_col, b_rpt, b_col, c_rpt, 0);
}

/* n_cols > 0: a row cannot hold more than n_cols distinct keys, so the table is sized from
 * min(products, n_cols) (keeps the heavy R-MAT rows' tables cache-sized; same counts). */
long long orc_spgemm_symbolic_n(int row_begin, int row_end, const int *a_rpt, const int *a_col, const int *b_rpt,
                                const int *b_col, long long *c_rpt, int n_cols)
{
    const int nrows = row_end - row_begin;
#pragma omp parallel
    {
        int *tab = NULL;
        size_t cap = 0;
#pragma omp for schedule(dynamic, 1)
        for (int ii = 0; ii < nrows; ii++) {
            const int i = row_begin + ii;
            long long ip = 0;
            for (int j = a_rpt[i]; j < a_rpt[i + 1]; j++) ip += b_rpt[a_col[j] + 1] - b_rpt[a_col[j]];
            if (n_cols > 0 && ip > n_cols) ip = n_cols;
            const size_t size = pow2_at_least(2 * ip);
            if (size > cap) {
                free(tab);
                tab = (int *)malloc(sizeof(int) * size);
                cap = size;
            }
            memset(tab, 0xff, sizeof(int) * size); /* init_check: -1 */
            long long nz = 0;
            for (int j = a_rpt[i]; j < a_rpt[i + 1]; j++) {
                const int acol = a_col[j];
                for (int k = b_rpt[acol]; k < b_rpt[acol + 1]; k++) {
                    const int key = b_col[k];
                    size_t h = ((uint32_t)key * ORC_HASH_SCAL) & (size - 1);
                    while (1) {
                        if (tab[h] == key) break;
                        if (tab[h] == -1) {
                            tab[h] = key;
                            nz++;
                            break;
                        }
                        h = (h + 1) & (size - 1);
                    }
                }
            }
            c_rpt[ii + 1] = nz;
        }
        free(tab);
    }
    c_rpt[0] = 0;
    for (int ii = 0; ii < nrows; ii++) c_rpt[ii + 1] += c_rpt[ii];
    return c_rpt[nrows];
}

typedef struct {
    int col;
    double val;
} orc_pair;

static int orc_pair_cmp(const void *a, const void *b)
{
    const int ca = ((const orc_pair *)a)->col, cb = ((const orc_pair *)b)->col;
    return (ca > cb) - (ca < cb);
}

/* ------------------------------------------------------------------------------------------
 * NUMERIC: per row, hash-accumulate aval*bval keyed by the B column (kernel_spgemm_hash_d.cu:
 * 866-889), then emit the row sorted by column (:917-925).  acc_double = 0 accumulates in
 * `real` exactly like the reference's atomic_fadd on `real` (sequential product order);
 * acc_double = 1 accumulates in double and rounds once at the end -- the tighter comparison
 * target used by the parity tests for fp32.
 * ------------------------------------------------------------------------------------------ */
#define ORC_DEFINE_NUMERIC(NAME, REAL)                                                                   \
    int NAME(int row_begin, int row_end, const int *a_rpt, const int *a_col, const REAL *a_val,         \
             const int *b_rpt, const int *b_col, const REAL *b_val, const long long *c_rpt, int *c_col, \
             REAL *c_val, int acc_double)                                                               \
    {                                                                                                   \
        const int nrows = row_end - row_begin;                                                          \
        int bad = 0;                                                                                    \
        _Pragma("omp parallel")                                                                         \
        {                                                                                               \
            int *keys = NULL;                                                                           \
            double *dacc = NULL;                                                                        \
            REAL *racc = NULL;                                                                          \
            orc_pair *out = NULL;                                                                       \
            size_t cap = 0, ocap = 0;                                                                   \
            _Pragma("omp for schedule(dynamic, 1)")                                                    \
            for (int ii = 0; ii < nrows; ii++) {                                                        \
                const int i = row_begin + ii;                                                           \
                const long long nnz = c_rpt[ii + 1] - c_rpt[ii];                                        \
                const size_t size = pow2_at_least(2 * nnz);                                             \
                if (size > cap) {                                                                       \
                    free(keys); free(dacc); free(racc);                                                 \
                    keys = (int *)malloc(sizeof(int) * size);                                           \
                    dacc = (double *)malloc(sizeof(double) * size);                                     \
                    racc = (REAL *)malloc(sizeof(REAL) * size);                                         \
                    cap = size;                                                                         \
                }                                                                                       \
                if ((size_t)nnz > ocap) {                                                               \
                    free(out);                                                                          \
                    out = (orc_pair *)malloc(sizeof(orc_pair) * (size_t)(nnz + 1));                     \
                    ocap = (size_t)nnz;                                                                 \
                }                                                                                       \
                memset(keys, 0xff, sizeof(int) * size);                                                 \
                for (size_t q = 0; q < size; q++) { dacc[q] = 0.0; racc[q] = (REAL)0; }                 \
                long long used = 0;                                                                     \
                for (int j = a_rpt[i]; j < a_rpt[i + 1]; j++) {                                         \
                    const int acol = a_col[j];                                                          \
                    const REAL aval = a_val[j];                                                         \
                    for (int k = b_rpt[acol]; k < b_rpt[acol + 1]; k++) {                               \
                        const int key = b_col[k];                                                       \
                        size_t h = ((uint32_t)key * ORC_HASH_SCAL) & (size - 1);                        \
                        while (1) {                                                                     \
                            if (keys[h] == key) break;                                                  \
                            if (keys[h] == -1) {                                                        \
                                if (used >= nnz) { bad = 1; break; }                                    \
                                keys[h] = key;                                                          \
                                used++;                                                                 \
                                break;                                                                  \
                            }                                                                           \
                            h = (h + 1) & (size - 1);                                                   \
                        }                                                                               \
                        if (keys[h] != key) continue;                                                   \
                        racc[h] += aval * b_val[k];                                                     \
                        dacc[h] += (double)aval * (double)b_val[k];                                     \
                    }                                                                                   \
                }                                                                                       \
                if (used != nnz) bad = 1;                                                               \
                long long n = 0;                                                                        \
                for (size_t q = 0; q < size && n < nnz; q++) {                                          \
                    if (keys[q] != -1) {                                                                \
                        out[n].col = keys[q];                                                           \
                        out[n].val = acc_double ? dacc[q] : (double)racc[q];                            \
                        n++;                                                                            \
                    }                                                                                   \
                }                                                                                       \
                qsort(out, (size_t)n, sizeof(orc_pair), orc_pair_cmp);                                  \
                for (long long q = 0; q < n; q++) {                                                     \
                    c_col[c_rpt[ii] + q] = out[q].col;                                                  \
                    c_val[c_rpt[ii] + q] = (REAL)out[q].val;                                            \
                }                                                                                       \
            }                                                                                           \
            free(keys); free(dacc); free(racc); free(out);                                              \
        }                                                                                               \
        return bad ? -1 : 0;                                                                            \
    }

ORC_DEFINE_NUMERIC(orc_spgemm_numeric_s, float)
ORC_DEFINE_NUMERIC(orc_spgemm_numeric_d, double)

/* ------------------------------------------------------------------------------------------
 * CPU SpMV, nsparse.cu:240-259: y[i] = sum_j val[j] * x[col[j]], accumulated left to right
 * in `real`.  Serial like the reference; orc_spmv_csr_*_omp is the row-parallel variant used
 * as the all-core baseline (identical results: rows are independent).
 * ------------------------------------------------------------------------------------------ */
#define ORC_DEFINE_SPMV(NAME, REAL, PAR)                                                           \
    void NAME(int M, const int *rpt, const int *col, const REAL *val, const REAL *x, REAL *y)      \
    {                                                                                              \
        PAR                                                                                        \
        for (int i = 0; i < M; i++) {                                                              \
            REAL ans = 0;                                                                          \
            for (int j = 0; j < rpt[i + 1] - rpt[i]; j++) ans += val[rpt[i] + j] * x[col[rpt[i] + j]]; \
            y[i] = ans;                                                                            \
        }                                                                                          \
    }

ORC_DEFINE_SPMV(orc_spmv_csr_s, float, )
ORC_DEFINE_SPMV(orc_spmv_csr_d, double, )
ORC_DEFINE_SPMV(orc_spmv_csr_s_omp, float, _Pragma("omp parallel for schedule(static)"))
ORC_DEFINE_SPMV(orc_spmv_csr_d_omp, double, _Pragma("omp parallel for schedule(static)"))
