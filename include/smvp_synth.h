/*
 * include/smvp_synth.h -- device-side synthetic matrix generators and COO shard helpers of
 * libsmvp_cuda.  Benchmark / test infrastructure for the configurations BASELINE.json names that
 * are too large to exist as Matrix Market files (the reference has no generator; its only input
 * path is the .mtx loader, main-cli.c:1405-1441).  Everything is counter-based (splitmix64 of the
 * coordinates / edge index), so any shard of any matrix can be generated independently and
 * reproducibly on any rank.
 *
 * Arrays returned through pointer-to-pointer arguments are device memory owned by the caller:
 * release with smvp_device_free().  COO is structure-of-arrays, 0-based, int32 + fp64.
 */
#ifndef SMVP_SYNTH_H
#define SMVP_SYNTH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

#define SMVP_VAL_STENCIL 0 /* 26.0 on the diagonal, -1.0 off it (27-point Laplacian-like)        */
#define SMVP_VAL_HASH 1    /* uniform(-1,1) hashed from (seed, global row, col)                   */
#define SMVP_VAL_ONES 2    /* 1.0 (pattern matrix, like the reference's pattern files)            */

/* 27-point stencil on an nx*ny*nz grid, row = x + nx*(y + ny*z), neighbours clipped at the faces.
 * Emits the rows [row_begin,row_end) in (row,col)-sorted order with LOCAL row indices
 * (row - row_begin) and global column indices.                                                   */
int smvp_synth_stencil27(int32_t nx, int32_t ny, int32_t nz, int64_t row_begin, int64_t row_end, int value_mode,
                         uint64_t seed, int32_t **d_row, int32_t **d_col, double **d_val, int64_t *nnz);
/* number of nonzeros in rows [0,row) of that stencil matrix (closed form, host only) */
int64_t smvp_synth_stencil27_prefix(int32_t nx, int32_t ny, int32_t nz, int64_t row);

/* R-MAT (Chakrabarti et al.): `nedges` draws on a 2^scale x 2^scale matrix with quadrant
 * probabilities (a,b,c,1-a-b-c), duplicates removed, result (row,col)-sorted.                    */
int smvp_synth_rmat(int scale, int64_t nedges, double a, double b, double c, int value_mode, uint64_t seed,
                    int32_t **d_row, int32_t **d_col, double **d_val, int64_t *nnz);

/* x[i] = uniform(-1,1) hashed from (seed, i)  -- or 1.0 when seed == 0 (the reference's ones vector) */
int smvp_synth_vector(double *d_x, int64_t n, uint64_t seed, void *stream);

/* entries with row in [row_lo,row_hi) and col in [col_lo,col_hi), order preserved, indices shifted by
 * -row_shift / -col_shift.  Used to cut row blocks (CSR) and column blocks (TJDS) for multi-GPU.   */
int smvp_coo_filter_device(const int32_t *d_row, const int32_t *d_col, const double *d_val, int64_t nnz,
                           int32_t row_lo, int32_t row_hi, int32_t col_lo, int32_t col_hi, int32_t row_shift,
                           int32_t col_shift, int32_t **o_row, int32_t **o_col, double **o_val, int64_t *o_nnz);
/* counts[k] = entries whose key (row if by_col == 0, else col) equals k; d_counts has nkeys uint32 */
int smvp_coo_histogram_device(const int32_t *d_row, const int32_t *d_col, int64_t nnz, int by_col, int32_t nkeys,
                              uint32_t *d_counts);

/* y[i] += a[i] over n doubles (fixed-order combine of partial results) */
int smvp_vector_add_device(double *d_y, const double *d_a, int64_t n, void *stream);

/* asynchronous device-to-device copy on `stream` (copy engines, no SM).  dst may be a peer mapping or an NVSwitch
 * multicast mapping of symmetric memory: one copy to the multicast address lands in every rank's buffer. */
int smvp_copy_device(void *d_dst, const void *d_src, int64_t bytes, void *stream);

/* the same copy done by a small SM kernel (`ctas` CTAs, 128-bit loads and stores) instead of the copy engines: useful
 * when dst is a multicast mapping, where coalesced SM stores move data faster than a copy-engine transfer */
int smvp_push_device(void *d_dst, const void *d_src, int64_t bytes, int ctas, void *stream);

/* the same with n_dst (<= 8) destinations: the source is read once and stored to every destination (the unicast
 * all-gather of one rank's block of y into its peers' buffers).  d_dst_list is a HOST array of device pointers. */
int smvp_push_fanout_device(void *const *d_dst_list, int n_dst, const void *d_src, int64_t bytes, int ctas, void *stream);

/* the same fan-out issued through the TMA engine: `ctas` one-warp CTAs stream 16 KB chunks of the source through shared
 * memory with cp.async.bulk and send every chunk to each destination with bulk stores -- full-size write packets, the
 * source read once, almost no SM time.  Falls back to smvp_push_fanout_device for mutually misaligned pointers. */
int smvp_push_tma_device(void *const *d_dst_list, int n_dst, const void *d_src, int64_t bytes, int ctas, void *stream);

/* out[i] = (((p_0[i] + p_1[i]) + p_2[i]) + ...) with p_k = d_parts + k * stride: the fixed-order combine of per-rank
 * partial results (column-block TJDS), bit-identical whatever order the parts arrived in */
int smvp_sum_ordered_device(double *d_out, const double *d_parts, int nparts, int64_t stride, int64_t n, void *stream);

/* the same with the parts given as nparts (<= 16) separate 16-byte-aligned device pointers (HOST array), e.g. the peer
 * mappings of a symmetric-memory buffer: the owner of a row block pulls that block of every rank's partial y over
 * NVLink and adds them in rank order */
int smvp_sum_ordered_ptrs_device(double *d_out, const double *const *d_part_list, int nparts, int64_t n, void *stream);

/* L2 flush helper for timing hygiene: writes `bytes` of a scratch buffer owned by the library */
int smvp_flush_l2(int64_t bytes, void *stream);

void smvp_device_free(void *d_ptr);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif
