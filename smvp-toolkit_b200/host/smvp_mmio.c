/*
 * smvp_mmio.c -- see smvp_mmio.h.  Banner and size-line parsing per the Matrix Market exchange format:
 *   line 1:  %%MatrixMarket <object> <format> <field> <symmetry>      (case-insensitive after the tag)
 *   then any number of lines starting with '%', optional blank lines, then "rows cols nnz".
 * Behaviour the reference relies on (and tests/test_host.py pins): an empty file yields
 * MM_PREMATURE_EOF (reference fixture sample-data/badfile.mtx -> main-cli.c:146-150), a first line
 * with fewer than five tokens yields MM_PREMATURE_EOF, a wrong tag MM_NO_HEADER, unknown words
 * MM_UNSUPPORTED_TYPE.
 */
#include "smvp_mmio.h"

#include <ctype.h>
#include <string.h>

static void lower(char *s)
{
    for (; *s; s++)
        *s = (char)tolower((unsigned char)*s);
}

int mm_is_valid(MM_typecode t)
{
    if (!mm_is_matrix(t))
        return 0;
    if (mm_is_dense(t) && mm_is_pattern(t))
        return 0;
    if (mm_is_real(t) && mm_is_hermitian(t))
        return 0;
    if (mm_is_pattern(t) && (mm_is_hermitian(t) || mm_is_skew(t)))
        return 0;
    return 1;
}

int mm_read_banner(FILE *f, MM_typecode *matcode)
{
    char line[MM_MAX_LINE_LENGTH];
    char tok[5][MM_MAX_TOKEN_LENGTH];
    int n = 0;
    char *p;

    mm_clear_typecode(matcode);
    if (!fgets(line, sizeof line, f))
        return MM_PREMATURE_EOF;

    /* split into at most five whitespace-separated words */
    p = line;
    while (n < 5)
    {
        size_t len;
        while (*p && isspace((unsigned char)*p))
            p++;
        if (!*p)
            break;
        len = strcspn(p, " \t\r\n\v\f");
        if (len >= MM_MAX_TOKEN_LENGTH)
            len = MM_MAX_TOKEN_LENGTH - 1;
        memcpy(tok[n], p, len);
        tok[n][len] = '\0';
        p += strcspn(p, " \t\r\n\v\f");
        n++;
    }
    if (n != 5)
        return MM_PREMATURE_EOF;
    if (strncmp(tok[0], MatrixMarketBanner, strlen(MatrixMarketBanner)) != 0)
        return MM_NO_HEADER;
    for (n = 1; n < 5; n++)
        lower(tok[n]);

    if (strcmp(tok[1], "matrix") != 0)
        return MM_UNSUPPORTED_TYPE;
    mm_set_matrix(matcode);

    if (strcmp(tok[2], "coordinate") == 0)
        mm_set_coordinate(matcode);
    else if (strcmp(tok[2], "array") == 0)
        mm_set_array(matcode);
    else
        return MM_UNSUPPORTED_TYPE;

    if (strcmp(tok[3], "real") == 0)
        mm_set_real(matcode);
    else if (strcmp(tok[3], "complex") == 0)
        mm_set_complex(matcode);
    else if (strcmp(tok[3], "pattern") == 0)
        mm_set_pattern(matcode);
    else if (strcmp(tok[3], "integer") == 0)
        mm_set_integer(matcode);
    else
        return MM_UNSUPPORTED_TYPE;

    if (strcmp(tok[4], "general") == 0)
        mm_set_general(matcode);
    else if (strcmp(tok[4], "symmetric") == 0)
        mm_set_symmetric(matcode);
    else if (strcmp(tok[4], "hermitian") == 0)
        mm_set_hermitian(matcode);
    else if (strcmp(tok[4], "skew-symmetric") == 0)
        mm_set_skew(matcode);
    else
        return MM_UNSUPPORTED_TYPE;
    return 0;
}

int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz)
{
    char line[MM_MAX_LINE_LENGTH];
    *M = *N = *nz = 0;
    for (;;)
    {
        if (!fgets(line, sizeof line, f))
            return MM_PREMATURE_EOF;
        if (line[0] == '%')
            continue; /* comment */
        if (sscanf(line, "%d %d %d", M, N, nz) == 3)
            return 0;
        /* blank or partial line: keep looking */
    }
}

int mm_write_banner(FILE *f, MM_typecode matcode)
{
    return fprintf(f, "%s %s\n", MatrixMarketBanner, mm_typecode_to_str(matcode)) < 0 ? MM_COULD_NOT_WRITE_FILE : 0;
}

int mm_write_mtx_crd_size(FILE *f, int M, int N, int nz)
{
    return fprintf(f, "%d %d %d\n", M, N, nz) < 0 ? MM_COULD_NOT_WRITE_FILE : 0;
}

char *mm_typecode_to_str(MM_typecode t)
{
    static char buf[4 * MM_MAX_TOKEN_LENGTH];
    const char *obj = mm_is_matrix(t) ? "matrix" : "?";
    const char *fmt = mm_is_sparse(t) ? "coordinate" : (mm_is_dense(t) ? "array" : "?");
    const char *fld = mm_is_real(t) ? "real" : mm_is_complex(t) ? "complex" : mm_is_pattern(t) ? "pattern" : mm_is_integer(t) ? "integer" : "?";
    const char *sym = mm_is_general(t) ? "general" : mm_is_symmetric(t) ? "symmetric" : mm_is_hermitian(t) ? "hermitian" : mm_is_skew(t) ? "skew-symmetric" : "?";
    snprintf(buf, sizeof buf, "%s %s %s %s", obj, fmt, fld, sym);
    return buf;
}
