/*
 * main-cli.c -- smvp-toolkit-cli on the B200 engine.
 *
 * Drop-in for the reference's command line (main-cli.c:1219-1481): same options, same messages, same
 * report files; the two algorithm bodies (smvp_csr_compute :325, smvp_tjds_compute :734) are replaced
 * by calls into libsmvp_cuda (include/smvp_cuda.h).
 *
 *   -a, --all-algs        enable all SMVP algorithms (CSR then TJDS)
 *   -c, --csr             enable CSR            -t, --tjds   enable TJDS
 *   -g, --cisr-gen        generate the CISR .coe image on stdout (host-side packing, smvp_cisr.c)
 *   -n, --number=INT      iterations per algorithm (default 1000)
 *   -s, --slots=INT       CISR slots (default 16)
 *   -d, --dir=FOLDER      report folder (default: current directory)
 *   <file>                Matrix Market file; options come BEFORE it (POSIXLY_CORRECT parsing, as popt's
 *                         POPT_CONTEXT_POSIXMEHARDER at main-cli.c:1254)
 * Additions (long options only, so no reference option changes meaning):
 *   --csr-variant=auto|vector|merge      --tjds-variant=atomic|deterministic|fast
 *   --ref-compat          TJDS walks only the diagonals the SHIPPED reference loop walks (main-cli.c:865,
 *                         :1013), reproducing its golden TJDS reports; default is the full product
 *   --json                print one machine-readable line per algorithm (GB/s, GFLOP/s, variant)
 *   --expand-symmetric    mirror the stored triangle of symmetric / skew-symmetric files (the reference multiplies
 *                         the stored triangle only; off by default to keep report parity)
 *
 * Deliberate differences from the reference at HEAD:
 *   U1  --all-algs runs CSR and TJDS (at HEAD its mask 256 matches no algorithm bit and nothing runs);
 *   U2  without -d the report goes to the current directory (HEAD reads an uninitialised pointer);
 *   U12 the COO list lives on the heap (HEAD: a stack VLA, ~524k entries at most);
 *   U13/debug dumps are gone (HEAD prints every array and 730 400 Verilog lines to stdout);
 *   U14 the ones vector has `cols` entries.
 */
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include "smvp_host.h"

#define ANSI_COLOR_RED "\x1b[31m"
#define ANSI_COLOR_GREEN "\x1b[32m"
#define ANSI_COLOR_YELLOW "\x1b[33m"
#define ANSI_COLOR_MAGENTA "\x1b[35m"
#define ANSI_COLOR_CYAN "\x1b[36m"
#define ANSI_COLOR_RESET "\x1b[0m"

/* the reference's algorithm bit mask (main-cli.c:34-38), with ALG_ALL actually covering the algorithms */
#define ALG_NONE 0
#define ALG_CSR (1 << 1)
#define ALG_TJDS (1 << 2)
#define ALG_CISR (1 << 3)
#define ALG_ALL (ALG_CSR | ALG_TJDS)

static void die(const char *msg)
{
    printf(ANSI_COLOR_RED "[ERROR]\t%s\n" ANSI_COLOR_RESET, msg);
    exit(1);
}

static void usage(const char *argv0)
{
    fprintf(stderr,
            "Usage: %s [-acgt?] [-a|--all-algs] [-c|--csr] [-g|--cisr-gen] [-t|--tjds] [-n|--number=1000]\n"
            "        [-s|--slots=16] [-d|--dir=./] [--csr-variant=auto|vector|merge]\n"
            "        [--tjds-variant=atomic|deterministic|fast] [--ref-compat] [--json] [--expand-symmetric] [-?|--help] [--usage] [OPTIONS] <file>\n",
            argv0);
}

static int folder_exists(const char *path) /* checkFolderExists, main-cli.c:1205-1217 */
{
    struct stat st;
    return stat(path, &st) == 0 && S_ISDIR(st.st_mode);
}

static void cuda_die(const char *what, int rc)
{
    printf(ANSI_COLOR_RED "[ERROR]\t%s failed: %s" ANSI_COLOR_RESET, what, smvp_strerror(rc));
    if (rc == SMVP_E_CUDA)
        printf(ANSI_COLOR_RED " (%s)" ANSI_COLOR_RESET, smvp_last_cuda_error());
    printf("\n");
    exit(1);
}

static int parse_int(const char *s, int *out)
{
    char *end = NULL;
    long v = strtol(s, &end, 10);
    if (end == s || *end != '\0')
        return 0;
    *out = (int)v;
    return 1;
}

int main(int argc, char *argv[])
{
    static const struct option longopts[] = {
        {"all-algs", no_argument, NULL, 'a'},    {"csr", no_argument, NULL, 'c'},
        {"cisr-gen", no_argument, NULL, 'g'},    {"tjds", no_argument, NULL, 't'},
        {"number", required_argument, NULL, 'n'}, {"slots", required_argument, NULL, 's'},
        {"dir", required_argument, NULL, 'd'},   {"help", no_argument, NULL, '?'},
        {"usage", no_argument, NULL, '?'},       {"csr-variant", required_argument, NULL, 1001},
        {"tjds-variant", required_argument, NULL, 1002}, {"ref-compat", no_argument, NULL, 1003},
        {"json", no_argument, NULL, 1004},       {"expand-symmetric", no_argument, NULL, 1005},
        {NULL, 0, NULL, 0}};
    int alg_mode = ALG_NONE, calc_iter = 1000, cisr_slots = 16; /* defaults of main-cli.c:1258-1264 */
    int csr_variant = SMVP_CSR_AUTO, tjds_variant = SMVP_TJDS_ATOMIC, ref_compat = 0, json = 0, all = 0, expand = 0;
    const char *report_dir = "";
    const char *input;
    int c, rows = 0, cols = 0, rc, i;
    int64_t nnz = 0;
    MM_typecode matcode;
    smvp_coo *coo = NULL;
    double *x, *y, *ms;
    int pinned = 1;
    smvp_time_stats_t st;
    char path[4096];

    if (argc < 2) /* main-cli.c:1267-1271 */
    {
        usage(argv[0]);
        return 1;
    }
    opterr = 0;
    while ((c = getopt_long(argc, argv, "+:acgtn:s:d:?", longopts, NULL)) != -1)
    {
        switch (c)
        {
        case 'a':
            if (alg_mode != ALG_NONE) /* main-cli.c:1279-1283 */
                die("Combining [-a|--all] with other algorithm flags is not supported.");
            alg_mode = ALG_ALL;
            all = 1;
            break;
        case 'c':
        case 't':
        case 'g':
            if (all) /* main-cli.c:1290-1321 */
                die("Combining [-a|--all] with other algorithm flags is not supported.");
            alg_mode |= (c == 'c') ? ALG_CSR : (c == 't') ? ALG_TJDS : ALG_CISR;
            break;
        case 'n':
            if (!parse_int(optarg, &calc_iter))
                die("Argument for iteration count contains non-number characters.");
            if (calc_iter < 1)
                die("Invalid number of algorithm iterations specified.");
            break;
        case 's':
            if (!parse_int(optarg, &cisr_slots) || cisr_slots < 1)
                die("Invalid number of CISR slots specified.");
            break;
        case 'd':
            if (!folder_exists(optarg))
                die("Report output folder not found. Check path and/or create folder if it does not exist.");
            report_dir = optarg;
            break;
        case 1001:
            if (strcmp(optarg, "auto") == 0)
                csr_variant = SMVP_CSR_AUTO;
            else if (strcmp(optarg, "vector") == 0)
                csr_variant = SMVP_CSR_VECTOR;
            else if (strcmp(optarg, "merge") == 0)
                csr_variant = SMVP_CSR_MERGE;
            else
                die("Unknown --csr-variant (auto, vector, merge).");
            break;
        case 1002:
            if (strcmp(optarg, "atomic") == 0)
                tjds_variant = SMVP_TJDS_ATOMIC;
            else if (strcmp(optarg, "deterministic") == 0)
                tjds_variant = SMVP_TJDS_DETERMINISTIC;
            else if (strcmp(optarg, "fast") == 0)
                tjds_variant = SMVP_TJDS_DETERMINISTIC_FAST;
            else
                die("Unknown --tjds-variant (atomic, deterministic, fast).");
            break;
        case 1003:
            ref_compat = 1;
            break;
        case 1004:
            json = 1;
            break;
        case 1005:
            expand = 1;
            break;
        case ':':
            die("One or more options missing a required argument.");
            break;
        default:
            usage(argv[0]);
            return 1;
        }
    }
    if (optind != argc - 1) /* exactly one positional (main-cli.c:1389-1393) */
    {
        usage(argv[0]);
        fprintf(stderr, ANSI_COLOR_RED "[ERROR]\tMust specify a single input file: ex., /path/to/file.mtx\n" ANSI_COLOR_RESET);
        return 1;
    }
    input = argv[optind];
    {
        FILE *probe = fopen(input, "r");
        if (!probe)
            die("Specified input file not found."); /* main-cli.c:1394-1398 */
        fclose(probe);
    }

    printf(ANSI_COLOR_GREEN "\n[START]\tExecuting smvp-toolbox-cli v%d.%d.%d\n" ANSI_COLOR_RESET, SMVP_MAJOR_VER, SMVP_MINOR_VER,
           SMVP_REVISION_VER);
    rc = smvp_load_mtx_ex(input, expand, &matcode, &rows, &cols, &nnz, &coo);
    if (rc != 0) /* mmioErrorHandler (main-cli.c:144-166) + the "only sparse" check (:1410-1414) */
        die(smvp_mmio_error_text(rc));
    printf(ANSI_COLOR_MAGENTA "[FILE]\tInput matrix file name: " ANSI_COLOR_RESET "%s\n", input);
    printf(ANSI_COLOR_YELLOW "[INFO]\tLoading matrix content from source file.\n" ANSI_COLOR_RESET);
    printf(ANSI_COLOR_CYAN "[DATA]\tNon-zero numbers contained in matrix: " ANSI_COLOR_RESET "%lld\n", (long long)nnz);
    printf(ANSI_COLOR_CYAN "[DATA]\tVector operand in use: " ANSI_COLOR_RESET "Ones vector with dimensions [%d, %d]\n", cols, 1);

    /* page-locked: smvp_csr_mult overlaps the transfers of big vectors with the multiply only from such buffers */
    x = (double *)smvp_host_alloc((int64_t)sizeof(double) * (cols > 0 ? cols : 1));
    y = (double *)smvp_host_alloc((int64_t)sizeof(double) * (rows > 0 ? rows : 1));
    if (!x || !y) /* no device / no page-locked memory: pageable buffers, the engine reports what is wrong */
    {
        smvp_host_free(x);
        smvp_host_free(y);
        pinned = 0;
        x = (double *)malloc(sizeof(double) * (size_t)(cols > 0 ? cols : 1));
        y = (double *)malloc(sizeof(double) * (size_t)(rows > 0 ? rows : 1));
    }
    ms = (double *)malloc(sizeof(double) * (size_t)calc_iter);
    if (!x || !y || !ms)
        die("Out of memory.");
    for (i = 0; i < cols; i++)
        x[i] = 1.0; /* vectorInit(.., 1), main-cli.c:368-369 */

    if (alg_mode & ALG_CSR)
    {
        smvp_csr *A = NULL;
        smvp_csr_info_t info;
        printf(ANSI_COLOR_YELLOW "[INFO]\tConverting loaded content to CSR format.\n" ANSI_COLOR_RESET);
        rc = smvp_csr_build(coo, rows, cols, nnz, &A);
        if (rc != SMVP_OK)
            cuda_die("smvp_csr_build", rc);
        printf(ANSI_COLOR_YELLOW "[INFO]\tCalculating %d iterations of SMVP CSR.\n" ANSI_COLOR_RESET, calc_iter);
        rc = smvp_csr_mult(A, x, y, calc_iter, ms, csr_variant);
        if (rc != SMVP_OK)
            cuda_die("smvp_csr_mult", rc);
        smvp_time_stats(ms, calc_iter, &st);
        smvp_csr_info(A, &info);
        if (smvp_write_report(input, report_dir, "CSR", (int)nnz, rows, calc_iter, y, &st, (unsigned long)time(NULL), path,
                              sizeof path) != 0)
            die("Could not write the report file.");
        printf(ANSI_COLOR_MAGENTA "[FILE]\tExecution report file saved as:\n" ANSI_COLOR_RESET);
        printf("\t%s\n", path);
        if (json)
            printf("{\"alg\": \"CSR\", \"variant\": \"%s\", \"rows\": %d, \"cols\": %d, \"nnz\": %lld, \"iters\": %d, \"avg_ms\": %.9g, "
                   "\"min_ms\": %.9g, \"gbps\": %.6g, \"gflops\": %.6g}\n",
                   (csr_variant == SMVP_CSR_AUTO ? info.auto_variant : csr_variant) == SMVP_CSR_VECTOR ? "vector" : "merge", rows, cols,
                   (long long)nnz, calc_iter, st.time_avg, st.time_min, info.bytes_per_mult / (st.time_avg * 1e6),
                   2.0 * nnz / (st.time_avg * 1e6));
        smvp_csr_free(A);
    }
    if (alg_mode & ALG_TJDS)
    {
        smvp_tjds *T = NULL;
        smvp_tjds_info_t info;
        printf(ANSI_COLOR_YELLOW "[INFO]\tConverting loaded content to TJDS format.\n" ANSI_COLOR_RESET);
        rc = smvp_tjds_build(coo, rows, cols, nnz, &T);
        if (rc != SMVP_OK)
            cuda_die("smvp_tjds_build", rc);
        smvp_tjds_info(T, &info);
        printf(ANSI_COLOR_YELLOW "[INFO]\tCalculating %d iterations of SMVP TJDS.\n" ANSI_COLOR_RESET, calc_iter);
        rc = smvp_tjds_mult(T, x, y, calc_iter, ms, tjds_variant, ref_compat ? info.ref_diag_limit : 0);
        if (rc != SMVP_OK)
            cuda_die("smvp_tjds_mult", rc);
        smvp_time_stats(ms, calc_iter, &st);
        if (smvp_write_report(input, report_dir, "TJDS", (int)nnz, rows, calc_iter, y, &st, (unsigned long)time(NULL), path,
                              sizeof path) != 0)
            die("Could not write the report file.");
        printf(ANSI_COLOR_MAGENTA "[FILE]\tExecution report file saved as:\n" ANSI_COLOR_RESET);
        printf("\t%s\n", path);
        if (json)
            printf("{\"alg\": \"TJDS\", \"variant\": \"%s\", \"rows\": %d, \"cols\": %d, \"nnz\": %lld, \"ndiag\": %d, \"iters\": %d, "
                   "\"avg_ms\": %.9g, \"min_ms\": %.9g, \"gbps\": %.6g, \"gflops\": %.6g, \"ref_compat\": %d}\n",
                   tjds_variant == SMVP_TJDS_ATOMIC ? "atomic" : (tjds_variant == SMVP_TJDS_DETERMINISTIC ? "deterministic" : "deterministic_fast"), rows, cols, (long long)nnz, info.ndiag, calc_iter,
                   st.time_avg, st.time_min, info.bytes_per_mult / (st.time_avg * 1e6), 2.0 * nnz / (st.time_avg * 1e6), ref_compat);
        smvp_tjds_free(T);
    }
    if (alg_mode & ALG_CISR)
    {
        /* smvp_cisr_coegen (main-cli.c:473-729): CSR through the engine's builder, packing on the host */
        smvp_csr *A = NULL;
        int32_t *row_ptr = (int32_t *)malloc(sizeof(int32_t) * ((size_t)rows + 1));
        int32_t *col_ind = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
        double *val = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
        if (!row_ptr || !col_ind || !val)
            die("Out of memory.");
        printf(ANSI_COLOR_YELLOW "[INFO]\tConverting loaded content to CISR format.\n" ANSI_COLOR_RESET);
        rc = smvp_csr_build(coo, rows, cols, nnz, &A);
        if (rc != SMVP_OK)
            cuda_die("smvp_csr_build", rc);
        rc = smvp_csr_export(A, row_ptr, col_ind, val);
        if (rc != SMVP_OK)
            cuda_die("smvp_csr_export", rc);
        smvp_csr_free(A);
        if (smvp_cisr_coe(stdout, row_ptr, col_ind, val, rows, nnz, cisr_slots) != 0)
            die("CISR COE generation failed (slot schedule overran the matrix).");
        free(row_ptr);
        free(col_ind);
        free(val);
    }

    printf(ANSI_COLOR_GREEN "[STOP]\tExit smvp-toolbox v%d.%d.%d\n\n" ANSI_COLOR_RESET, SMVP_MAJOR_VER, SMVP_MINOR_VER,
           SMVP_REVISION_VER);
    if (pinned)
    {
        smvp_host_free(x);
        smvp_host_free(y);
    }
    else
    {
        free(x);
        free(y);
    }
    free(ms);
    free(coo);
    return 0;
}
