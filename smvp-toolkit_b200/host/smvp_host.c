/* smvp_host.c -- loader, report writer and messages of the host side.  See smvp_host.h. */
#include "smvp_host.h"

#include <ctype.h>
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

const char *smvp_mmio_error_text(int code)
{
    /* the four messages of mmioErrorHandler (main-cli.c:146-164), verbatim text */
    switch (code)
    {
    case MM_PREMATURE_EOF:
        return "Could not process specified Matrix Market input file. Required parameters not present on first line of file.";
    case MM_NO_HEADER:
        return "Could not process specified Matrix Market input file. Required header is missing or file contents may not be "
               "Matrix Market formatted.";
    case MM_UNSUPPORTED_TYPE:
        return "Could not process specified Matrix Market input file. Matrix content description not parseable or is absent.";
    case SMVP_HOST_E_OPEN:
        return "Specified input file not found.";
    case SMVP_HOST_E_NOT_SPARSE:
        return "This application only supports sparse matricies. Specified input file does not appear to contain a sparse matrix.";
    case SMVP_HOST_E_COMPLEX:
        return "Could not process specified Matrix Market input file. Complex-valued matrices are not supported.";
    case SMVP_HOST_E_ENTRIES:
        return "Could not process specified Matrix Market input file. Fewer (or malformed) entries than the size line declares.";
    case SMVP_HOST_E_ALLOC:
        return "Out of memory while loading the matrix.";
    default:
        return "Could not process specified Matrix Market input file. Unhandled exception occured during file loading .";
    }
}

/* one slice of a regular file, read with pread by its own thread (page-cache copies scale with threads) */
typedef struct
{
    int fd;
    char *dst;
    off_t off;
    size_t len;
    int ok;
} read_slice;

static void *read_slice_fn(void *arg)
{
    read_slice *r = (read_slice *)arg;
    size_t done = 0;
    r->ok = 1;
    while (done < r->len)
    {
        const ssize_t got = pread(r->fd, r->dst + done, r->len - done, r->off + (off_t)done);
        if (got <= 0)
        {
            r->ok = 0;
            break;
        }
        done += (size_t)got;
    }
    return NULL;
}

static int load_threads(size_t len);

/* read the rest of the stream into one NUL-terminated buffer */
static char *slurp(FILE *f, size_t *len)
{
    size_t cap = 1 << 20, n = 0;
    char *buf;
    struct stat st;
    long pos = ftell(f);
    if (pos >= 0 && fstat(fileno(f), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > pos)
    {
        /* regular file: one allocation, and the bytes are fetched by a few threads at once */
        const size_t total = (size_t)(st.st_size - pos);
        const int nt = load_threads(total);
        cap = total + 1;
        if (nt > 1 && (buf = (char *)malloc(cap + 1)) != NULL)
        {
            read_slice rs[256];
            pthread_t tid[256];
            int started[256], k, ok = 1;
            for (k = 0; k < nt; k++)
            {
                const size_t a = total / (size_t)nt * (size_t)k, b = k + 1 == nt ? total : total / (size_t)nt * (size_t)(k + 1);
                rs[k].fd = fileno(f);
                rs[k].dst = buf + a;
                rs[k].off = (off_t)pos + (off_t)a;
                rs[k].len = b - a;
                rs[k].ok = 0;
            }
            for (k = 1; k < nt; k++)
                started[k] = pthread_create(&tid[k], NULL, read_slice_fn, &rs[k]) == 0;
            read_slice_fn(&rs[0]);
            for (k = 1; k < nt; k++)
            {
                if (started[k])
                    pthread_join(tid[k], NULL);
                else
                    read_slice_fn(&rs[k]);
            }
            for (k = 0; k < nt; k++)
                ok = ok && rs[k].ok;
            if (ok)
            {
                buf[total] = '\0';
                *len = total;
                return buf;
            }
            free(buf); /* a short read (file shrank?): fall back to the plain stream read below */
            if (fseek(f, pos, SEEK_SET) != 0)
                return NULL;
        }
    }
    buf = (char *)malloc(cap + 1);
    if (!buf)
        return NULL;
    for (;;)
    {
        size_t got = fread(buf + n, 1, cap - n, f);
        n += got;
        if (got == 0)
            break;
        if (n == cap)
        {
            char *nb;
            cap *= 2;
            nb = (char *)realloc(buf, cap + 1);
            if (!nb)
            {
                free(buf);
                return NULL;
            }
            buf = nb;
        }
    }
    buf[n] = '\0';
    *len = n;
    return buf;
}

/* ------------------------------------------------------------------ parallel entry parser
 * The reference reads one entry per fscanf("%d %d %lg") (main-cli.c:1427-1441), about 10^6 entries/s: hopeless for the
 * GB-scale files the engine is built for.  The buffer is cut into line-aligned chunks, one per thread:
 * pass 1 counts the tokens of every chunk, pass 2 parses every chunk straight into its slice of the output.  The
 * semantics stay those of the sequential token parser below (which stays the reference for them): free-form white
 * space, exactly nnz entries read, anything after them ignored.  Whenever a chunk does not hold a whole number of
 * entries (an entry split over two lines at a chunk border) or a token is not a plain number, the parallel path gives
 * up and the sequential parser decides -- so every accepted file yields bit-identical entries either way.
 * Values: digits/exponent within Clinger's exact range (<= 15 significant digits, |exp10| <= 22) are converted with one
 * correctly rounded multiply or divide, which is what strtod returns; everything else goes through strtod itself. */
typedef struct
{
    const char *begin, *end; /* chunk [begin, end) */
    int pattern;
    int64_t tokens;          /* pass 1 */
    int64_t first_entry;     /* global index of the chunk's first entry */
    int64_t max_entries;     /* entries this chunk may write (global cap nnz) */
    smvp_coo *out;
    int bad;                 /* pass 2: a token the fast grammar does not cover */
} load_chunk;

static int is_ws(unsigned char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

static void *count_tokens(void *arg)
{
    load_chunk *c = (load_chunk *)arg;
    const char *p = c->begin;
    int64_t n = 0;
    int in_tok = 0;
    for (; p < c->end; p++)
    {
        const int ws = is_ws((unsigned char)*p);
        n += (!ws && !in_tok);
        in_tok = !ws;
    }
    c->tokens = n;
    return NULL;
}

/* decimal integer token -> long; returns the position after it, or NULL when the token is anything else */
static const char *fast_int(const char *p, const char *end, long *out)
{
    int neg = 0, nd = 0;
    unsigned long v = 0;
    if (p < end && (*p == '+' || *p == '-'))
        neg = *p++ == '-';
    while (p < end && *p >= '0' && *p <= '9' && nd < 18)
    {
        v = v * 10 + (unsigned long)(*p++ - '0');
        nd++;
    }
    if (nd == 0 || (p < end && !is_ws((unsigned char)*p)))
        return NULL;
    *out = neg ? -(long)v : (long)v;
    return p;
}

static const double POW10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

/* floating-point token -> double, bit-identical to strtod; NULL when the token is not a number strtod takes whole */
static const char *fast_double(const char *p, const char *end, double *out)
{
    const char *tok = p, *q;
    int neg = 0, nd = 0, nfrac = 0, seen_dot = 0, any = 0;
    unsigned long long m = 0;
    long e10 = 0;
    char *se;
    if (p < end && (*p == '+' || *p == '-'))
        neg = *p++ == '-';
    for (; p < end; p++)
    {
        if (*p >= '0' && *p <= '9')
        {
            any = 1;
            if (nd > 0 || *p != '0') /* leading zeros are not significant */
            {
                if (nd < 19)
                    m = m * 10 + (unsigned long long)(*p - '0');
                nd++;
            }
            nfrac += seen_dot;
        }
        else if (*p == '.' && !seen_dot)
            seen_dot = 1;
        else
            break;
    }
    if (any && p < end && (*p == 'e' || *p == 'E' || *p == 'd' || *p == 'D'))
    {
        /* 'd' exponents are Fortran's; strtod stops before them, so let strtod (and the sequential parser) judge */
        if (*p == 'd' || *p == 'D')
            any = 0;
        else
        {
            int eneg = 0, ed = 0;
            long ev = 0;
            q = p + 1;
            if (q < end && (*q == '+' || *q == '-'))
                eneg = *q++ == '-';
            while (q < end && *q >= '0' && *q <= '9' && ed < 6)
            {
                ev = ev * 10 + (*q++ - '0');
                ed++;
            }
            if (ed == 0)
                any = 0;
            else
            {
                e10 = eneg ? -ev : ev;
                p = q;
            }
        }
    }
    if (any && (p == end || is_ws((unsigned char)*p)) && nd <= 15)
    {
        const long e = e10 - nfrac;
        double v = (double)m; /* exact: m < 10^15 < 2^53 */
        if (m == 0)
        {
            *out = neg ? -0.0 : 0.0;
            return p;
        }
        if (e >= 0 && e <= 22)
        {
            v *= POW10[e];
            *out = neg ? -v : v;
            return p;
        }
        if (e < 0 && e >= -22)
        {
            v /= POW10[-e];
            *out = neg ? -v : v;
            return p;
        }
    }
    /* long mantissas, big exponents, inf / nan / hex: strtod decides, and it must consume the whole token */
    *out = strtod(tok, &se);
    if (se == tok || (se < end && !is_ws((unsigned char)*se)))
        return NULL;
    return se;
}

static void *parse_chunk(void *arg)
{
    load_chunk *c = (load_chunk *)arg;
    const char *p = c->begin;
    int64_t i;
    for (i = 0; i < c->max_entries; i++)
    {
        long r, col;
        double v = 1.0; /* pattern => 1 (main-cli.c:1432) */
        while (p < c->end && is_ws((unsigned char)*p))
            p++;
        if (!(p = fast_int(p, c->end, &r)))
            break;
        while (p < c->end && is_ws((unsigned char)*p))
            p++;
        if (!(p = fast_int(p, c->end, &col)))
            break;
        if (!c->pattern)
        {
            while (p < c->end && is_ws((unsigned char)*p))
                p++;
            if (!(p = fast_double(p, c->end, &v)))
                break;
        }
        c->out[i].row = (int32_t)(r - 1);
        c->out[i].col = (int32_t)(col - 1);
        c->out[i].val = v;
    }
    c->bad = i != c->max_entries;
    return NULL;
}

static int load_threads(size_t len)
{
    const char *e = getenv("SMVP_LOAD_THREADS");
    long n = e && e[0] ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    const char *m = getenv("SMVP_LOAD_MIN_CHUNK"); /* bytes of text per thread below which threads are not worth it */
    const long min_chunk = m && m[0] && atol(m) > 0 ? atol(m) : (4L << 20);
    const long by_size = (long)(len / (size_t)min_chunk) + 1;
    if (n > by_size)
        n = by_size;
    if (n > 256)
        n = 256;
    return n < 1 ? 1 : (int)n;
}

static void run_chunks(load_chunk *ch, int n, void *(*fn)(void *))
{
    pthread_t tid[256];
    int started[256], k;
    for (k = 1; k < n; k++)
        started[k] = pthread_create(&tid[k], NULL, fn, &ch[k]) == 0;
    fn(&ch[0]);
    for (k = 1; k < n; k++)
    {
        if (started[k])
            pthread_join(tid[k], NULL);
        else
            fn(&ch[k]); /* could not start a thread: do its share here */
    }
}

/* 1 = out[0 .. nz) filled; 0 = not applicable, the sequential parser must run */
static int parse_parallel(const char *buf, size_t len, int pattern, int64_t nz, smvp_coo *out)
{
    load_chunk ch[256];
    const int per_entry = pattern ? 2 : 3;
    int n = load_threads(len), k;
    int64_t entries = 0;
    const char *p = buf, *end = buf + len;
    if (n < 2 || nz == 0)
        return 0;
    for (k = 0; k < n; k++)
    {
        const char *stop = k + 1 == n ? end : buf + (len / (size_t)n) * (size_t)(k + 1);
        if (stop < p)
            stop = p;
        if (k + 1 < n)
        {
            const char *nl = (const char *)memchr(stop, '\n', (size_t)(end - stop));
            stop = nl ? nl + 1 : end;
        }
        ch[k].begin = p;
        ch[k].end = stop;
        ch[k].pattern = pattern;
        ch[k].bad = 0;
        p = stop;
    }
    run_chunks(ch, n, count_tokens);
    for (k = 0; k < n; k++)
    {
        int64_t e;
        if (ch[k].tokens % per_entry != 0 && entries + ch[k].tokens / per_entry < nz)
            return 0; /* an entry straddles a chunk border before the nz-th one */
        e = ch[k].tokens / per_entry;
        ch[k].first_entry = entries;
        ch[k].max_entries = entries >= nz ? 0 : (entries + e > nz ? nz - entries : e);
        ch[k].out = out + (entries < nz ? entries : nz);
        entries += e;
    }
    if (entries < nz)
        return 0; /* too few entries: let the sequential parser produce the error */
    run_chunks(ch, n, parse_chunk);
    for (k = 0; k < n; k++)
        if (ch[k].bad)
            return 0;
    return 1;
}

int smvp_load_mtx(const char *path, MM_typecode *matcode, int *rows, int *cols, int64_t *nnz, smvp_coo **coo)
{
    return smvp_load_mtx_ex(path, 0, matcode, rows, cols, nnz, coo);
}

int smvp_load_mtx_ex(const char *path, int expand_symmetric, MM_typecode *matcode, int *rows, int *cols, int64_t *nnz,
                     smvp_coo **coo)
{
    FILE *f;
    int rc, nz = 0;
    char *buf, *p, *end;
    size_t len = 0;
    smvp_coo *out;
    int64_t i;

    *coo = NULL;
    *rows = *cols = 0;
    *nnz = 0;
    f = fopen(path, "r");
    if (!f)
        return SMVP_HOST_E_OPEN;
    rc = mm_read_banner(f, matcode); /* main-cli.c:1405 */
    if (rc != 0)
    {
        fclose(f);
        return rc;
    }
    if (!mm_is_sparse(*matcode)) /* main-cli.c:1410 */
    {
        fclose(f);
        return SMVP_HOST_E_NOT_SPARSE;
    }
    if (mm_is_complex(*matcode))
    {
        fclose(f);
        return SMVP_HOST_E_COMPLEX;
    }
    rc = mm_read_mtx_crd_size(f, rows, cols, &nz); /* main-cli.c:1419 */
    if (rc != 0)
    {
        fclose(f);
        return rc;
    }
    if (nz < 0 || *rows < 0 || *cols < 0)
    {
        fclose(f);
        return SMVP_HOST_E_ENTRIES;
    }
    buf = slurp(f, &len);
    fclose(f);
    if (!buf)
        return SMVP_HOST_E_ALLOC;
    out = (smvp_coo *)malloc(sizeof(smvp_coo) * (size_t)(nz > 0 ? nz : 1) * (expand_symmetric ? 2 : 1));
    if (!out)
    {
        free(buf);
        return SMVP_HOST_E_ALLOC;
    }
    /* main-cli.c:1427-1441: "%d %d\n" for pattern files, "%d %d %lg\n" otherwise; then -1 on both indices */
    p = buf;
    i = 0;
    if (parse_parallel(buf, len, mm_is_pattern(*matcode) ? 1 : 0, nz, out))
        i = nz; /* all entries parsed by the chunked path; the loop below is skipped */
    for (; i < nz; i++)
    {
        long r, c;
        double v = 1.0; /* pattern => 1 (main-cli.c:1432) */
        errno = 0;
        r = strtol(p, &end, 10);
        if (end == p)
            break;
        p = end;
        c = strtol(p, &end, 10);
        if (end == p)
            break;
        p = end;
        if (!mm_is_pattern(*matcode))
        {
            v = strtod(p, &end);
            if (end == p)
                break;
            p = end;
        }
        out[i].row = (int32_t)(r - 1);
        out[i].col = (int32_t)(c - 1);
        out[i].val = v;
    }
    free(buf);
    if (i != nz)
    {
        free(out);
        return SMVP_HOST_E_ENTRIES;
    }
    *nnz = nz;
    if (expand_symmetric && !mm_is_general(*matcode))
    {
        const double sign = mm_is_skew(*matcode) ? -1.0 : 1.0;
        int64_t k = nz;
        for (i = 0; i < nz; i++)
            if (out[i].row != out[i].col)
            {
                out[k].row = out[i].col;
                out[k].col = out[i].row;
                out[k].val = sign * out[i].val;
                k++;
            }
        *nnz = k;
    }
    *coo = out;
    return 0;
}

/* The vector part of a report, "%g" per row (main-cli.c:309-314), for vectors of millions of rows: every thread
 * formats a contiguous block of rows with the same snprintf("%g") into its own buffer, the buffers are written in
 * order.  Same bytes as the fprintf loop; -1 = could not (memory / threads), the caller then runs that loop. */
#define REPORT_PARALLEL_ROWS (1 << 18)

typedef struct
{
    const double *y;
    int begin, end, rows;
    char *buf;
    size_t len;
} fmt_block;

static void *fmt_block_fn(void *arg)
{
    fmt_block *b = (fmt_block *)arg;
    size_t cap = (size_t)(b->end - b->begin) * 26 + 8, n = 0; /* "%g" of a double is at most 13 characters */
    int i;
    b->buf = (char *)malloc(cap);
    b->len = 0;
    if (!b->buf)
        return NULL;
    for (i = b->begin; i < b->end; i++)
    {
        n += (size_t)snprintf(b->buf + n, cap - n, "%g", b->y[i]);
        n += (size_t)snprintf(b->buf + n, cap - n, i < b->rows - 1 ? "\n" : "\n]\n\n");
    }
    b->len = n;
    return NULL;
}

static int write_vector_parallel(FILE *f, const double *y, int rows)
{
    fmt_block blk[256];
    int n = load_threads((size_t)rows * 64), k, rc = 0;
    if (n < 2)
        return -1;
    for (k = 0; k < n; k++)
    {
        blk[k].y = y;
        blk[k].rows = rows;
        blk[k].begin = (int)((int64_t)rows * k / n);
        blk[k].end = (int)((int64_t)rows * (k + 1) / n);
        blk[k].buf = NULL;
    }
    {
        pthread_t tid[256];
        int started[256];
        for (k = 1; k < n; k++)
            started[k] = pthread_create(&tid[k], NULL, fmt_block_fn, &blk[k]) == 0;
        fmt_block_fn(&blk[0]);
        for (k = 1; k < n; k++)
        {
            if (started[k])
                pthread_join(tid[k], NULL);
            else
                fmt_block_fn(&blk[k]);
        }
    }
    for (k = 0; k < n; k++)
        if (!blk[k].buf)
            rc = -1;
    for (k = 0; k < n && rc == 0; k++)
        if (fwrite(blk[k].buf, 1, blk[k].len, f) != blk[k].len)
            rc = -2; /* partial output: do not let the caller append the vector a second time */
    for (k = 0; k < n; k++)
        free(blk[k].buf);
    return rc; /* -1: nothing written, the caller may format sequentially; -2: a short write (disk full) -- an error */
}

int smvp_write_report(const char *input_file_name, const char *report_dir, const char *alg_name, int nnz, int rows,
                      int iters, const double *y, const smvp_time_stats_t *t, unsigned long unix_time, char *out_path,
                      size_t out_path_len)
{
    char name[128];
    char *full;
    size_t dlen = report_dir ? strlen(report_dir) : 0;
    FILE *f;
    int i;

    snprintf(name, sizeof name, "smvp-toolbox_report_%s_%lu.txt", alg_name, unix_time); /* main-cli.c:269 */
    full = (char *)malloc(dlen + sizeof name + 2);
    if (!full)
        return -1;
    if (dlen == 0)
        strcpy(full, name); /* main-cli.c:272-276 */
    else
    {
        strcpy(full, report_dir); /* main-cli.c:279-286 */
        if (report_dir[dlen - 1] != '/')
            strcat(full, "/");
        strcat(full, name);
    }
    if (out_path && out_path_len > 0)
    {
        strncpy(out_path, full, out_path_len - 1);
        out_path[out_path_len - 1] = '\0';
    }
    f = fopen(full, "a+"); /* main-cli.c:293 */
    free(full);
    if (!f)
        return -1;
    /* main-cli.c:294-316, field for field */
    fprintf(f, "Execution results for smvp-toolbox v.%d.%d.%d, %s algorithm\n", SMVP_MAJOR_VER, SMVP_MINOR_VER,
            SMVP_REVISION_VER, alg_name);
    fprintf(f, "Generated on %lu (Unix time)\n\n", unix_time);
    fprintf(f, "Sparse matrix file in use:\n%s\n\n", input_file_name);
    fprintf(f, "Non-zero numbers contained in matrix: %d\n\n", nnz);
    fprintf(f, "Compute times for %d iterations:\n\n", iters);
    fprintf(f, "Total Time: %g ms\n", t->time_total);
    fprintf(f, "Average Time: %g ms\n", t->time_avg);
    fprintf(f, "Fastest Time: %g ms\n", t->time_min);
    fprintf(f, "Slowest Time: %g ms\n", t->time_max);
    fprintf(f, "Time StDev: %g ms\n\n", t->time_stdev);
    fprintf(f, "Output vector (one cell per line):\n");
    fprintf(f, "[\n");
    if (rows >= REPORT_PARALLEL_ROWS)
    {
        const int prc = write_vector_parallel(f, y, rows);
        if (prc == 0 || prc == -2)
        {
            /* a short write, a failed flush or a failed close (ENOSPC, EIO) is an error: never report "saved" then */
            const int bad = prc == -2 || ferror(f);
            return (fclose(f) != 0 || bad) ? -1 : 0;
        }
    }
    for (i = 0; i < rows; i++)
    {
        fprintf(f, "%g", y[i]);
        fprintf(f, i < rows - 1 ? "\n" : "\n]\n\n");
    }
    {
        const int bad = ferror(f);
        return (fclose(f) != 0 || bad) ? -1 : 0;
    }
}
