/*
 * oracle/smvp_oracle.c -- TEST INFRASTRUCTURE ONLY (never shipped, never timed as the product).
 *
 * CPU restatement, in plain C, of the reference's CSR and TJDS hot path
 * (circletile/smvp-toolkit, /root/reference/main-cli.c).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's shared object.  The product (libsmvp_cuda) never does.
 *
 * PARITY PINNING: this restatement is pinned against
 *   (1) every golden report the reference ships for the path
 *       (output-test/smvp-toolbox_report_{CSR,TJDS}_*.txt, build/..._CSR_1619162887.txt),
 *       copied to tests/golden/reports/ and checked by tests/test_oracle_golden.py;
 *   (2) the known-answer arrays for pdp08-pg4 the reference prints with its debug
 *       switches;
 *   (3) the UNMODIFIED reference translation unit compiled into oracle/_ref/ and run
 *       in the build container (tests/golden/gen_golden.py, tests/test_oracle_vs_ref.py).
 *
 * Every function cites the reference lines it follows.  Where the reference has
 * undefined behaviour the intended value is used and the site is named (SURVEY.md
 * appendix A, U1..U15); a `ref_compat` mode reproduces the shipped TJDS truncation so
 * the golden TJDS reports can be matched digit for digit.
 *
 * Build: gcc -O3 -ffp-contract=off -fPIC -shared  (no FMA contraction: the reference
 * binary is plain x86-64 -O3, mulsd+addsd; main-cli.c:414, build/build.ninja:111).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>

/* == MMRawData, main-cli.c:42-47 (16 bytes, no padding) */
typedef struct
{
    int32_t row;
    int32_t col;
    double val;
} oracle_coo;

/* == MMDataPlus, main-cli.c:51-57: the reference overwrites `row` with the rank of the
 * entry inside its column and keeps the original row beside it. */
typedef struct
{
    int32_t rank;
    int32_t row_orig;
    int32_t col;
    double val;
} oracle_ranked;

/* == TXTable, main-cli.c:79-83.  `len` is (entries in column) - 1, as in the reference
 * (main-cli.c:851,859 store the LAST rank, not the count); empty columns get -1 here,
 * the reference leaves them uninitialised (U10). */
typedef struct
{
    int32_t origin_col;
    int32_t len;
} oracle_txrow;

/* ordering of main-cli.c:171-185 (row, then col) */
static int by_row_col(const void *a, const void *b)
{
    const oracle_coo *p = (const oracle_coo *)a, *q = (const oracle_coo *)b;
    if (p->row != q->row)
        return p->row < q->row ? -1 : 1;
    if (p->col != q->col)
        return p->col < q->col ? -1 : 1;
    return 0;
}

/* ordering of main-cli.c:190-204 (col, then row) */
static int by_col_row(const void *a, const void *b)
{
    const oracle_coo *p = (const oracle_coo *)a, *q = (const oracle_coo *)b;
    if (p->col != q->col)
        return p->col < q->col ? -1 : 1;
    if (p->row != q->row)
        return p->row < q->row ? -1 : 1;
    return 0;
}

/* ordering of main-cli.c:209-223 (length descending, then origin column ascending) */
static int by_len_desc(const void *a, const void *b)
{
    const oracle_txrow *p = (const oracle_txrow *)a, *q = (const oracle_txrow *)b;
    if (p->len != q->len)
        return p->len > q->len ? -1 : 1;
    if (p->origin_col != q->origin_col)
        return p->origin_col < q->origin_col ? -1 : 1;
    return 0;
}

/* ordering of main-cli.c:228-242 applied to MMDataPlus: (rank, then permuted col) */
static int by_rank_col(const void *a, const void *b)
{
    const oracle_ranked *p = (const oracle_ranked *)a, *q = (const oracle_ranked *)b;
    if (p->rank != q->rank)
        return p->rank < q->rank ? -1 : 1;
    if (p->col != q->col)
        return p->col < q->col ? -1 : 1;
    return 0;
}

/*
 * CSR build -- main-cli.c:340-365.
 *   :340      qsort by (row, col)  [the reference sorts the caller's array in place; we copy]
 *   :350-351  val / col_ind are the sorted sequence
 *   :353-364  row_ptr[r+1] = index of the first entry after row r; the else-if chain leaves
 *             row_ptr[0] and the slots of empty rows unwritten (U3).  Intended value: the
 *             exclusive scan of the row counts, which is what the written slots hold.
 * Returns 0, or -1 on allocation failure / out-of-range coordinates.
 */
int oracle_csr_build(const oracle_coo *coo, int32_t rows, int32_t cols, int64_t nnz,
                     int32_t *row_ptr, int32_t *col_ind, double *val)
{
    int64_t i;
    oracle_coo *s = (oracle_coo *)malloc(sizeof(oracle_coo) * (size_t)(nnz > 0 ? nnz : 1));
    if (!s)
        return -1;
    memcpy(s, coo, sizeof(oracle_coo) * (size_t)nnz);
    qsort(s, (size_t)nnz, sizeof(oracle_coo), by_row_col);

    for (i = 0; i <= rows; i++)
        row_ptr[i] = 0;
    for (i = 0; i < nnz; i++)
    {
        if (s[i].row < 0 || s[i].row >= rows || s[i].col < 0 || s[i].col >= cols)
        {
            free(s);
            return -1;
        }
        val[i] = s[i].val;
        col_ind[i] = s[i].col;
        row_ptr[s[i].row + 1] = (int32_t)(i + 1); /* last write per row wins == :357-360 / :353-356 */
    }
    /* fill the holes the reference leaves (empty rows, row 0): running maximum */
    for (i = 1; i <= rows; i++)
        if (row_ptr[i] < row_ptr[i - 1])
            row_ptr[i] = row_ptr[i - 1];
    free(s);
    return 0;
}

/*
 * CSR multiply -- main-cli.c:410-416, the loop body restated one to one:
 * y is zeroed by the caller-visible vectorInit at :405 (outside the timed bracket),
 * then y[i] += val[j] * x[col_ind[j]] left to right.
 */
void oracle_csr_mult(int32_t rows, const int32_t *row_ptr, const int32_t *col_ind,
                     const double *val, const double *x, double *y)
{
    int32_t r, j;
    for (r = 0; r < rows; r++)
        y[r] = 0.0;
    for (r = 0; r < rows; r++)
        for (j = row_ptr[r]; j < row_ptr[r + 1]; j++)
            y[r] += val[j] * x[col_ind[j]];
}

/*
 * Timed CSR loop for the CPU baseline: the reference's bracket (main-cli.c:402-420):
 * zero y untimed, CLOCK_MONOTONIC_RAW around the double loop only, ms per iteration.
 */
void oracle_csr_mult_timed(int32_t rows, const int32_t *row_ptr, const int32_t *col_ind,
                           const double *val, const double *x, double *y, int iters, double *ms_each)
{
    int it;
    int32_t r, j;
    struct timespec t0, t1;
    for (it = 0; it < iters; it++)
    {
        for (r = 0; r < rows; r++)
            y[r] = 0.0;
        clock_gettime(CLOCK_MONOTONIC_RAW, &t0);
        for (r = 0; r < rows; r++)
            for (j = row_ptr[r]; j < row_ptr[r + 1]; j++)
                y[r] += val[j] * x[col_ind[j]];
        clock_gettime(CLOCK_MONOTONIC_RAW, &t1);
        ms_each[it] = ((t1.tv_sec * 1e9 + t1.tv_nsec) - (t0.tv_sec * 1e9 + t0.tv_nsec)) / 1e6;
    }
}

/*
 * The same loop, rows cut into nnz-balanced blocks over `nthreads` host threads -- NOT something the reference
 * does (it is single-threaded); bench.py reports it next to the 1-thread number as "what every core of the box
 * could do with the reference's loop".  Each row is still summed left to right by one thread, so y is
 * bit-identical to oracle_csr_mult.  ms_each = wall time of each pass, first thread in to last thread out.
 */
typedef struct
{
    int32_t r0, r1;
    const int32_t *row_ptr, *col_ind;
    const double *val, *x;
    double *y;
    int iters;
    pthread_barrier_t *bar;
} csr_mt_arg;

static void *csr_mt_fn(void *argp)
{
    csr_mt_arg *a = (csr_mt_arg *)argp;
    int it;
    int32_t r, j;
    for (it = 0; it < a->iters; it++)
    {
        for (r = a->r0; r < a->r1; r++)
            a->y[r] = 0.0;
        pthread_barrier_wait(a->bar); /* pass starts */
        for (r = a->r0; r < a->r1; r++)
            for (j = a->row_ptr[r]; j < a->row_ptr[r + 1]; j++)
                a->y[r] += a->val[j] * a->x[a->col_ind[j]];
        pthread_barrier_wait(a->bar); /* pass ends */
    }
    return NULL;
}

int oracle_csr_mult_timed_mt(int32_t rows, const int32_t *row_ptr, const int32_t *col_ind, const double *val,
                             const double *x, double *y, int iters, double *ms_each, int nthreads)
{
    pthread_t *tid;
    csr_mt_arg *args;
    pthread_barrier_t bar;
    struct timespec t0, t1;
    int k, it, started = 0;
    const int64_t nnz = rows > 0 ? row_ptr[rows] : 0;
    if (nthreads < 1)
        nthreads = 1;
    tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    args = (csr_mt_arg *)malloc(sizeof(csr_mt_arg) * (size_t)nthreads);
    if (!tid || !args || pthread_barrier_init(&bar, NULL, (unsigned)nthreads + 1) != 0)
    {
        free(tid);
        free(args);
        return -1;
    }
    for (k = 0; k < nthreads; k++)
    {
        /* first row whose prefix reaches k/nthreads of the nonzeros (the engine's row-block rule) */
        int32_t lo = 0, hi = rows;
        const int64_t target = nnz * k / nthreads;
        while (lo < hi)
        {
            const int32_t mid = lo + (hi - lo) / 2;
            if (row_ptr[mid] < target)
                lo = mid + 1;
            else
                hi = mid;
        }
        args[k].r0 = k == 0 ? 0 : lo;
        if (k > 0)
            args[k - 1].r1 = args[k].r0;
        args[k].row_ptr = row_ptr;
        args[k].col_ind = col_ind;
        args[k].val = val;
        args[k].x = x;
        args[k].y = y;
        args[k].iters = iters;
        args[k].bar = &bar;
    }
    args[nthreads - 1].r1 = rows;
    for (k = 0; k < nthreads; k++)
    {
        if (pthread_create(&tid[k], NULL, csr_mt_fn, &args[k]) != 0)
            break;
        started++;
    }
    if (started != nthreads) /* cannot run short-handed: the barrier counts every thread */
    {
        for (k = 0; k < started; k++)
            pthread_cancel(tid[k]);
        for (k = 0; k < started; k++)
            pthread_join(tid[k], NULL);
        pthread_barrier_destroy(&bar);
        free(tid);
        free(args);
        return -1;
    }
    for (it = 0; it < iters; it++)
    {
        pthread_barrier_wait(&bar);
        clock_gettime(CLOCK_MONOTONIC_RAW, &t0);
        pthread_barrier_wait(&bar);
        clock_gettime(CLOCK_MONOTONIC_RAW, &t1);
        ms_each[it] = ((t1.tv_sec * 1e9 + t1.tv_nsec) - (t0.tv_sec * 1e9 + t0.tv_nsec)) / 1e6;
    }
    for (k = 0; k < nthreads; k++)
        pthread_join(tid[k], NULL);
    pthread_barrier_destroy(&bar);
    free(tid);
    free(args);
    return 0;
}

/*
 * TJDS build -- main-cli.c:766-967, phase by phase:
 *   :766      sort by (col, row)
 *   :789-826  rank of each entry inside its column (0 for the first of a column, +1 after);
 *             duplicates are undefined in the reference (U9) and rejected by callers
 *   :845-862  txList[c] = {c, last rank of column c}
 *   :865      num_tjdiag read BEFORE the sort (U4) -> returned as *ref_ndiag_limit
 *   :868      sort txList by (len desc, col asc)  -> perm[p] = origin column at slot p
 *   :894-904  col := slot of col (linear search there; inverse permutation here, same result)
 *   :926      sort by (rank, slot)
 *   :944-967  val / row_ind in that order; start_pos[d] = first index of rank d,
 *             terminal entry = nnz (unwritten in the reference when the last diagonal has one
 *             element, U5; always written here)
 * Outputs: perm[cols], start_pos[*ndiag + 1] (caller allocates rows+2 or cols... see below),
 *          row_ind[nnz], val[nnz], *ndiag = number of jagged diagonals (max column count),
 *          *ref_ndiag_limit = count(col 0) + 1, the number of diagonals the shipped loop walks
 *          (main-cli.c:865 with :1013 `index < num_tjdiag + 1`).
 * start_pos must have room for (max column count + 1) <= rows + 1 entries.
 */
int oracle_tjds_build(const oracle_coo *coo, int32_t rows, int32_t cols, int64_t nnz,
                      int32_t *perm, int32_t *start_pos, int32_t *ndiag, int32_t *ref_ndiag_limit,
                      int32_t *row_ind, double *val)
{
    int64_t i;
    int32_t c, d;
    int rc = -1;
    oracle_coo *s = (oracle_coo *)malloc(sizeof(oracle_coo) * (size_t)(nnz > 0 ? nnz : 1));
    oracle_ranked *e = (oracle_ranked *)malloc(sizeof(oracle_ranked) * (size_t)(nnz > 0 ? nnz : 1));
    oracle_txrow *tx = (oracle_txrow *)malloc(sizeof(oracle_txrow) * (size_t)(cols > 0 ? cols : 1));
    int32_t *slot_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cols > 0 ? cols : 1));
    if (!s || !e || !tx || !slot_of)
        goto done;

    memcpy(s, coo, sizeof(oracle_coo) * (size_t)nnz);
    qsort(s, (size_t)nnz, sizeof(oracle_coo), by_col_row); /* :766 */

    for (c = 0; c < cols; c++)
    {
        tx[c].origin_col = c;
        tx[c].len = -1; /* U10: empty column sorts last */
    }
    for (i = 0; i < nnz; i++) /* :789-826, :845-862 */
    {
        if (s[i].row < 0 || s[i].row >= rows || s[i].col < 0 || s[i].col >= cols)
            goto done;
        e[i].row_orig = s[i].row;
        e[i].col = s[i].col;
        e[i].val = s[i].val;
        e[i].rank = (i > 0 && s[i].col == s[i - 1].col) ? e[i - 1].rank + 1 : 0;
        tx[s[i].col].len = e[i].rank; /* last write per column wins == :848-853, :856-861 */
    }
    *ref_ndiag_limit = (cols > 0 ? tx[0].len + 1 : 0) + 1; /* :865 then the `+ 1` of :1013 */

    qsort(tx, (size_t)cols, sizeof(oracle_txrow), by_len_desc); /* :868 */
    for (c = 0; c < cols; c++)
    {
        perm[c] = tx[c].origin_col;
        slot_of[tx[c].origin_col] = c;
    }
    for (i = 0; i < nnz; i++) /* :894-904 */
        e[i].col = slot_of[e[i].col];

    qsort(e, (size_t)nnz, sizeof(oracle_ranked), by_rank_col); /* :926 */

    d = 0;
    for (i = 0; i < nnz; i++) /* :944-967 */
    {
        val[i] = e[i].val;
        row_ind[i] = e[i].row_orig;
        if (i == 0 || e[i].rank > e[i - 1].rank)
            start_pos[d++] = (int32_t)i;
    }
    start_pos[d] = (int32_t)nnz;
    *ndiag = d;
    rc = 0;
done:
    free(s);
    free(e);
    free(tx);
    free(slot_of);
    return rc;
}

/*
 * TJDS multiply, intended semantics of main-cli.c:1013-1020: every diagonal,
 * y[row_ind[j]] += val[j] * x_perm[j - start_pos[d]] with x_perm[p] = x[perm[p]]
 * (the reference permutes x once at build time, :907-923, and then indexes it by ROW, U7;
 * with its hard-wired x = ones the two are indistinguishable).
 * diag_limit <= 0: all diagonals.  diag_limit = k > 0: only the first min(k, ndiag).
 */
void oracle_tjds_mult(int32_t rows, int32_t cols, int32_t ndiag, const int32_t *perm,
                      const int32_t *start_pos, const int32_t *row_ind, const double *val,
                      const double *x, double *y, int32_t diag_limit)
{
    int32_t d, j, r, lim = (diag_limit > 0 && diag_limit < ndiag) ? diag_limit : ndiag;
    double *xp = (double *)malloc(sizeof(double) * (size_t)(cols > 0 ? cols : 1));
    for (j = 0; j < cols; j++)
        xp[j] = x[perm[j]];
    for (r = 0; r < rows; r++)
        y[r] = 0.0;
    for (d = 0; d < lim; d++)
        for (j = start_pos[d]; j < start_pos[d + 1]; j++)
        {
            r = row_ind[j];
            y[r] += val[j] * xp[j - start_pos[d]];
        }
    free(xp);
}

/*
 * TJDS multiply exactly as SHIPPED (for matching the golden TJDS report files only):
 *   - walks diagonals 0 .. ref_ndiag_limit-1 (main-cli.c:865 + :1013), clipped to ndiag;
 *     slots of start_pos past the last written one are treated as 0, i.e. an empty range,
 *     which is what the fresh verbatim run shows (malloc'd, untouched memory; U6);
 *   - drops the last diagonal when it holds exactly one element, because its terminal
 *     start_pos entry is never written (main-cli.c:957-966, U5);
 *   - indexes x by ROW (main-cli.c:1017-1018, U7).
 */
void oracle_tjds_mult_ref_compat(int32_t rows, int32_t ndiag, int32_t ref_ndiag_limit,
                                 const int32_t *start_pos, const int32_t *row_ind,
                                 const double *val, const double *x_by_row, double *y)
{
    int32_t d, j, r, lim = ref_ndiag_limit < ndiag ? ref_ndiag_limit : ndiag;
    for (r = 0; r < rows; r++)
        y[r] = 0.0;
    for (d = 0; d < lim; d++)
    {
        if (d == ndiag - 1 && start_pos[d + 1] - start_pos[d] == 1)
            break;
        for (j = start_pos[d]; j < start_pos[d + 1]; j++)
        {
            r = row_ind[j];
            y[r] += val[j] * x_by_row[r];
        }
    }
}

/* Timed TJDS loop for the CPU baseline (bracket of main-cli.c:1004-1024), full product. */
void oracle_tjds_mult_timed(int32_t rows, int32_t ndiag, const int32_t *start_pos,
                            const int32_t *row_ind, const double *val, const double *x_perm,
                            double *y, int iters, double *ms_each)
{
    int it;
    int32_t d, j, r;
    struct timespec t0, t1;
    for (it = 0; it < iters; it++)
    {
        for (r = 0; r < rows; r++)
            y[r] = 0.0;
        clock_gettime(CLOCK_MONOTONIC_RAW, &t0);
        for (d = 0; d < ndiag; d++)
            for (j = start_pos[d]; j < start_pos[d + 1]; j++)
            {
                r = row_ind[j];
                y[r] += val[j] * x_perm[j - start_pos[d]];
            }
        clock_gettime(CLOCK_MONOTONIC_RAW, &t1);
        ms_each[it] = ((t1.tv_sec * 1e9 + t1.tv_nsec) - (t0.tv_sec * 1e9 + t0.tv_nsec)) / 1e6;
    }
}

/*
 * Iteration statistics -- main-cli.c:428-456 and calcStDevDouble :114-130:
 * total, mean, min, max and the POPULATION standard deviation sqrt(sum((t-mean)^2)/n).
 * The reference never initialises its accumulators (U11); zero is the intended start.
 * out[5] = {total, avg, stdev, min, max}  (field order of struct _time_data_, :87-95).
 */
void oracle_time_stats(const double *ms_each, int n, double *out)
{
    int i;
    double total = 0.0, mn = 0.0, mx = 0.0, ss = 0.0, mean;
    for (i = 0; i < n; i++)
    {
        total += ms_each[i];
        if (i == 0 || ms_each[i] < mn)
            mn = ms_each[i];
        if (i == 0 || ms_each[i] > mx)
            mx = ms_each[i];
    }
    mean = n > 0 ? total / n : 0.0;
    for (i = 0; i < n; i++)
        ss += pow(ms_each[i] - mean, 2);
    out[0] = total;
    out[1] = mean;
    out[2] = n > 0 ? sqrt(ss / n) : 0.0;
    out[3] = mn;
    out[4] = mx;
}
