/* smvp_host.c -- loader, report writer and messages of the host side.  See smvp_host.h. */
#include "smvp_host.h"

#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

/* read the rest of the stream into one NUL-terminated buffer */
static char *slurp(FILE *f, size_t *len)
{
    size_t cap = 1 << 20, n = 0;
    char *buf = (char *)malloc(cap + 1);
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
    for (i = 0; i < nz; i++)
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
    for (i = 0; i < rows; i++)
    {
        fprintf(f, "%g", y[i]);
        fprintf(f, i < rows - 1 ? "\n" : "\n]\n\n");
    }
    fclose(f);
    return 0;
}
