/*
 * smvp_host.h -- the C host side around libsmvp_cuda: Matrix Market loader, report writer, messages.
 * These are the pieces of the reference's main-cli.c that sit either side of the hot path and that a
 * drop-in has to keep byte-compatible:
 *     loader   main-cli.c:1405-1441  (banner/size through mmio, then one fscanf per entry)
 *     report   generateReportText, main-cli.c:246-320
 *     errors   mmioErrorHandler, main-cli.c:144-166
 */
#ifndef SMVP_HOST_H
#define SMVP_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "smvp_mmio.h"
#include "smvp_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SMVP_MAJOR_VER 0 /* the reference's version triple (main-cli.c:7-9): it is part of the report header */
#define SMVP_MINOR_VER 6
#define SMVP_REVISION_VER 4

/* loader errors beyond the MM_* codes of smvp_mmio.h */
#define SMVP_HOST_E_OPEN 101      /* fopen failed                                                      */
#define SMVP_HOST_E_NOT_SPARSE 102 /* array (dense) file: "only supports sparse matricies" (:1410-1414) */
#define SMVP_HOST_E_COMPLEX 103   /* complex field: the reference mis-parses these (U15); rejected      */
#define SMVP_HOST_E_ENTRIES 104   /* fewer / malformed entries than the size line promises             */
#define SMVP_HOST_E_ALLOC 105

/*
 * Load a coordinate Matrix Market file with the reference's semantics (main-cli.c:1426-1441):
 * indices 1-based -> 0-based, `pattern` files get val = 1.0, integer/real values parsed as double,
 * symmetric / skew / hermitian files are NOT expanded (only the stored triangle is used, as the
 * reference does).  *coo is malloc'd (free() it).  Returns 0 or an MM_* / SMVP_HOST_E_* code.
 * Unlike the reference (a stack VLA, main-cli.c:1426, and one fscanf per entry) the entries live on
 * the heap, the file is fetched with parallel preads and parsed by one thread per line-aligned chunk
 * (SMVP_LOAD_THREADS, default: all online cores; 1 = the sequential token parser), so GB-scale files
 * load at tens of millions of entries per second.  The entries are bit-identical whatever the thread count.
 */
int smvp_load_mtx(const char *path, MM_typecode *matcode, int *rows, int *cols, int64_t *nnz, smvp_coo **coo);

/*
 * Same, with optional expansion of the stored triangle (the reference never expands,
 * main-cli.c:1427-1441, so its product on pwt.mtx is that of the lower triangle only):
 * expand_symmetric != 0 appends (col,row,val) for every off-diagonal entry of a `symmetric` file and
 * (col,row,-val) for a `skew-symmetric` one (`hermitian` real files are treated as symmetric).
 * Off by default everywhere so that report parity with the reference is preserved.
 */
int smvp_load_mtx_ex(const char *path, int expand_symmetric, MM_typecode *matcode, int *rows, int *cols, int64_t *nnz,
                     smvp_coo **coo);

/* the reference's message for an mmio error code (mmioErrorHandler, main-cli.c:144-166), without colour codes */
const char *smvp_mmio_error_text(int code);

/*
 * Write one report in the reference's exact format (main-cli.c:294-316) to
 *     <report_dir>/smvp-toolbox_report_<alg_name>_<unix_time>.txt        (opened "a+", main-cli.c:293)
 * report_dir NULL or "" means the current directory (the reference leaves the pointer uninitialised
 * without -d, U2).  The path written is returned in out_path when it is not NULL.
 */
int smvp_write_report(const char *input_file_name, const char *report_dir, const char *alg_name, int nnz, int rows,
                      int iters, const double *y, const smvp_time_stats_t *t, unsigned long unix_time, char *out_path,
                      size_t out_path_len);

/*
 * The reference's `-g` option (smvp_cisr_coegen, main-cli.c:473-729): CSR -> CISR slot schedule -> 36-bit
 * words of a Xilinx block-RAM .coe image, written to `out` (the reference prints to stdout).  Host-only.
 * Returns 0, -1 on bad arguments / allocation failure, -2 where the reference aborts ("slot_group_iter overran").
 */
int smvp_cisr_coe(FILE *out, const int32_t *row_ptr, const int32_t *col_ind, const double *val, int rows, int64_t nnz,
                  int slots);

#ifdef __cplusplus
}
#endif
#endif
