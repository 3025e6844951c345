// tjds_mult.cu -- y = A x in TJDS on sm_100a.  Replaces the loop at main-cli.c:1013-1020
//
//     for d in diagonals: for j in start_pos[d] .. start_pos[d+1]:  y[row_ind[j]] += val[j] * x_perm[j - start_pos[d]]
//
// (the reference indexes x by ROW there, U7; with its x = ones that is invisible -- the intended
// index, the slot of the entry's column, is used here).
//
// Layout fact the kernels exploit: slot p of EVERY diagonal belongs to the same (permuted) column, and
// columns are sorted by length, so "thread p walks diagonals 0 .. len[p]-1 at offset p" reads val and
// row_ind fully coalesced (consecutive threads, consecutive addresses, per diagonal), keeps x_perm[p]
// in a register, and lanes of a warp have near-equal trip counts.  One launch covers all diagonals
// (never one launch per diagonal: power-law matrices have 1e5..1e6 of them).
//
//  ATOMIC         scatter with fp64 atomicAdd (RED.E.ADD.F64) into y.
//  DETERMINISTIC  same walk, but every product is split exactly into two signed 64-bit fixed-point
//                 words (scaled per row by a bound known before the multiply) and accumulated with
//                 INTEGER atomics.  Integer addition is associative, so the result is bit-identical
//                 run to run whatever order the hardware schedules the adds in; a final pass converts
//                 the 96-bit sums to fp64 (one rounding, i.e. more accurate than any fp64 summation
//                 order).  Cost: two 8-byte reductions per entry instead of one, 16 B/row of
//                 accumulators to clear and read back.
#include "common.cuh"

namespace smvp
{

// SMVP_TJDS_STREAM=1 reads the matrix streams with ld.global.nc.L1::no_allocate; measured on B200 it LOSES to
// plain __ldg (atomic 3.19 vs 3.12 ms, deterministic 4.12 vs 3.83 ms on the 369^3 stencil), so it stays off
#ifndef SMVP_TJDS_STREAM
#define SMVP_TJDS_STREAM 0
#endif
__device__ __forceinline__ int32_t ld_stream_i32(const int32_t *p)
{
#if SMVP_TJDS_STREAM
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ double ld_stream_f64(const double *p)
{
#if SMVP_TJDS_STREAM
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

constexpr int TJDS_W = 32;           // bits kept in the low word
constexpr int TJDS_FRAC = 62 + TJDS_W; // value = V * 2^(T_r - TJDS_FRAC), |sum V| < 2^TJDS_FRAC

__global__ void __launch_bounds__(256) tjds_permute_x_kernel(const double *__restrict__ x, const int32_t *__restrict__ perm,
                                                             int32_t cols, double *__restrict__ x_perm, int32_t *__restrict__ x_exp)
{
    int32_t e = INT32_MIN;
    for (int32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < cols; p += gridDim.x * blockDim.x)
    {
        const double v = x[perm[p]];
        x_perm[p] = v;
        if (v != 0.0)
            e = max(e, ilogb(v) + 1); // |v| < 2^e
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        e = max(e, __shfl_xor_sync(0xffffffffu, e, o));
    if ((threadIdx.x & 31) == 0 && e != INT32_MIN)
        atomicMax(x_exp, e);
}

// ---- work plan: power-law matrices have columns with 1e5..1e6 entries; one thread walking such a
// column alone would serialise the whole multiply.  The (diagonal, slot) plane is therefore cut into
// segments of TJDS_SEG consecutive diagonals; segment g only needs the first L[g*TJDS_SEG] slots (the
// columns longer than g*TJDS_SEG), i.e. ceil(L/256) CTAs.  blocks[b] = {segment, first slot} is a flat
// table over all segments, built once per handle, so one launch covers everything and no thread walks
// more than TJDS_SEG entries.
constexpr int TJDS_SEG = 32;

__global__ void __launch_bounds__(256) tjds_seg_count_kernel(const int32_t *__restrict__ start_pos, int32_t ndiag, int32_t nseg,
                                                             uint32_t *__restrict__ seg_nblocks)
{
    const int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nseg)
        return;
    const int32_t d = g * TJDS_SEG;
    const int32_t L = start_pos[d + 1] - start_pos[d];
    seg_nblocks[g] = (uint32_t)((L + 255) / 256);
}

__global__ void __launch_bounds__(256) tjds_seg_fill_kernel(const uint32_t *__restrict__ seg_first, const int32_t *__restrict__ start_pos,
                                                            int32_t nseg, int2 *__restrict__ blocks)
{
    // one CTA per segment: writes that segment's run of the table
    const int32_t g = blockIdx.x;
    const int32_t d = g * TJDS_SEG;
    const int32_t L = start_pos[d + 1] - start_pos[d];
    const int32_t nb = (L + 255) / 256;
    const uint32_t first = seg_first[g];
    for (int32_t i = threadIdx.x; i < nb; i += blockDim.x)
        blocks[first + i] = make_int2(g, i * 256);
}

template <int UNROLL>
__global__ void __launch_bounds__(256) tjds_atomic_kernel(const int2 *__restrict__ blocks, const int32_t *__restrict__ start_pos,
                                                          const int32_t *__restrict__ slot_len, const int32_t *__restrict__ row_ind,
                                                          const double *__restrict__ val, const double *__restrict__ x_perm,
                                                          double *__restrict__ y, int32_t nslots, int32_t diag_limit)
{
    const int2 blk = __ldg(blocks + blockIdx.x);
    const int32_t p = blk.y + threadIdx.x;
    if (p >= nslots)
        return;
    const int32_t d_begin = blk.x * TJDS_SEG;
    const int32_t len = min(min(slot_len[p], diag_limit), d_begin + TJDS_SEG);
    const double xp = x_perm[p];
    int32_t d = d_begin;
    for (; d + UNROLL <= len; d += UNROLL)
    {
        int32_t r[UNROLL];
        double v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            const int64_t j = (int64_t)__ldg(start_pos + d + u) + p;
            r[u] = ld_stream_i32(row_ind + j);
            v[u] = ld_stream_f64(val + j);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
            atomicAdd(y + r[u], __dmul_rn(v[u], xp));
    }
    for (; d < len; d++)
    {
        const int64_t j = (int64_t)__ldg(start_pos + d) + p;
        atomicAdd(y + ld_stream_i32(row_ind + j), __dmul_rn(ld_stream_f64(val + j), xp));
    }
}

// ---- deterministic variant --------------------------------------------------------------------
// row_exp[r] = ea_r + cb_r where |a_rj| < 2^ea_r for every entry of row r and the row holds at most
// 2^cb_r entries.  With |x_c| < 2^ex every partial sum of row r is below 2^(row_exp[r] + ex) = 2^T_r.
__global__ void __launch_bounds__(256) tjds_row_bound_kernel(const int32_t *__restrict__ row_ind, const double *__restrict__ val,
                                                             int64_t nnz, int32_t *__restrict__ row_maxexp, uint32_t *__restrict__ row_cnt)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += (int64_t)gridDim.x * blockDim.x)
    {
        const int32_t r = row_ind[j];
        const double v = val[j];
        atomicAdd(row_cnt + r, 1u);
        if (v != 0.0)
            atomicMax(row_maxexp + r, ilogb(v) + 1);
    }
}

__global__ void __launch_bounds__(256) tjds_row_exp_kernel(int32_t *__restrict__ row_exp, const uint32_t *__restrict__ row_cnt, int32_t rows)
{
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows)
        return;
    const uint32_t c = row_cnt[r];
    const int32_t cb = c <= 1 ? 0 : 32 - __clz(c - 1); // ceil(log2(c))
    const int32_t e = row_exp[r];
    row_exp[r] = (e == EXP_NONE) ? EXP_NONE : e + cb;
}

// exact split of one fp64 product into (hi, lo) fixed-point words for a row with bound 2^T
__device__ __forceinline__ void fixed_split(double prod, int32_t T, long long *hi, long long *lo)
{
    *hi = 0;
    *lo = 0;
    if (prod == 0.0)
        return;
    const long long bits = __double_as_longlong(prod);
    const int32_t be = (int32_t)((bits >> 52) & 0x7ff);
    unsigned long long m = (unsigned long long)bits & 0xfffffffffffffULL;
    int32_t e; // prod = m * 2^e
    if (be == 0)
        e = -1074;
    else
    {
        m |= 1ULL << 52;
        e = be - 1075;
    }
    const int32_t s = e - (T - TJDS_FRAC); // V = m * 2^s, |V| < 2^TJDS_FRAC
    unsigned long long hm, lm;
    if (s >= TJDS_W)
    {
        hm = m << (s - TJDS_W);
        lm = 0;
    }
    else if (s >= 0)
    {
        hm = m >> (TJDS_W - s);
        lm = (m << s) & ((1ULL << TJDS_W) - 1ULL);
    }
    else if (s > -64)
    {
        const unsigned long long t = m >> (-s); // bits below 2^(T-94) are dropped (far below fp64 resolution of the row)
        hm = t >> TJDS_W;
        lm = t & ((1ULL << TJDS_W) - 1ULL);
    }
    else
    {
        hm = 0;
        lm = 0;
    }
    const bool neg = bits < 0;
    *hi = neg ? -(long long)hm : (long long)hm;
    *lo = neg ? -(long long)lm : (long long)lm;
}

template <int UNROLL>
__global__ void __launch_bounds__(256) tjds_det_kernel(const int2 *__restrict__ blocks, const int32_t *__restrict__ start_pos,
                                                       const int32_t *__restrict__ slot_len, const int32_t *__restrict__ row_ind,
                                                       const double *__restrict__ val, const double *__restrict__ x_perm,
                                                       const int32_t *__restrict__ row_exp, const int32_t *__restrict__ x_exp,
                                                       unsigned long long *__restrict__ acc, int32_t rows, int32_t nslots,
                                                       int32_t diag_limit)
{
    const int2 blk = __ldg(blocks + blockIdx.x);
    const int32_t p = blk.y + threadIdx.x;
    if (p >= nslots)
        return;
    const int32_t d_begin = blk.x * TJDS_SEG;
    const int32_t len = min(min(slot_len[p], diag_limit), d_begin + TJDS_SEG);
    const double xp = x_perm[p];
    const int32_t ex = __ldg(x_exp);
    for (int32_t d0 = d_begin; d0 < len; d0 += UNROLL)
    {
        int32_t r[UNROLL];
        double v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            r[u] = -1;
            v[u] = 0.0;
            if (d0 + u < len)
            {
                const int64_t j = (int64_t)__ldg(start_pos + d0 + u) + p;
                r[u] = ld_stream_i32(row_ind + j);
                v[u] = ld_stream_f64(val + j);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            if (r[u] >= 0)
            {
                long long hi, lo;
                fixed_split(__dmul_rn(v[u], xp), __ldg(row_exp + r[u]) + ex, &hi, &lo);
                if (hi != 0)
                    atomicAdd(acc + r[u], (unsigned long long)hi); // hi words [0, rows), lo words [rows, 2 rows):
                if (lo != 0)
                    atomicAdd(acc + (int64_t)rows + r[u], (unsigned long long)lo); // a warp's reductions stay contiguous
            }
        }
    }
}

// row_rank != NULL: the accumulators and row_exp are in popularity-rank order (relabelled handle), y is not
__global__ void __launch_bounds__(256) tjds_det_finalize_kernel(const long long *__restrict__ acc, const int32_t *__restrict__ row_exp,
                                                                const int32_t *__restrict__ x_exp, int32_t rows, double *__restrict__ y,
                                                                const int32_t *__restrict__ row_rank)
{
    const int32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows)
        return;
    const int32_t r = row_rank ? row_rank[row] : row;
    longlong2 a;
    a.x = acc[r];
    a.y = acc[(int64_t)rows + r];
    double out = 0.0;
    if (a.x != 0 || a.y != 0)
    {
        // (hi * 2^W + lo) as one correctly-ordered fp64 sum, then one exact scaling by a power of two
        const double s = __dadd_rn(__dmul_rn((double)a.x, 4294967296.0), (double)a.y);
        const int32_t T = row_exp[r] + *x_exp;
        out = scalbn(s, T - TJDS_FRAC);
    }
    y[row] = out;
}

// y[r] = y_rel[rank[r]]: the atomic variant of a relabelled handle sums in rank order
__global__ void __launch_bounds__(256) tjds_unrank_y_kernel(const double *__restrict__ y_rel, const int32_t *__restrict__ row_rank,
                                                            int32_t rows, double *__restrict__ y)
{
    for (int32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x)
        y[r] = __ldg(y_rel + __ldg(row_rank + r));
}

// the row indices the kernels scatter through: the relabelled copy when that plan is in use (relabel.cu)
static inline const int32_t *mult_rows(const smvp_tjds *A) { return A->relabel_state == 1 ? A->row_rel : A->row_ind; }

static int tjds_plan(smvp_tjds *A, cudaStream_t s)
{
    if (A->seg_blocks || A->ndiag == 0)
        return SMVP_OK;
    const int32_t nseg = (A->ndiag + TJDS_SEG - 1) / TJDS_SEG;
    uint32_t *seg_nb = nullptr, *d_total = nullptr;
    SMVP_CUDA(dev_alloc(&seg_nb, nseg));
    SMVP_CUDA(dev_alloc(&d_total, 1));
    SMVP_LAUNCH(tjds_seg_count_kernel, (unsigned)ceil_div64(nseg, 256), 256, 0, s, (const int32_t *)A->start_pos, A->ndiag, nseg, seg_nb);
    SMVP_TRY(exclusive_scan_u32(seg_nb, seg_nb, nseg, d_total, s));
    uint32_t total = 0;
    SMVP_CUDA(cudaMemcpyAsync(&total, d_total, sizeof(total), cudaMemcpyDeviceToHost, s));
    SMVP_CUDA(cudaStreamSynchronize(s));
    SMVP_CUDA(dev_alloc(&A->seg_blocks, total));
    A->num_seg_blocks = (int32_t)total;
    SMVP_LAUNCH(tjds_seg_fill_kernel, (unsigned)nseg, 256, 0, s, (const uint32_t *)seg_nb, (const int32_t *)A->start_pos, nseg,
                A->seg_blocks);
    SMVP_CUDA(cudaStreamSynchronize(s));
    SMVP_CUDA(cudaFree(seg_nb));
    SMVP_CUDA(cudaFree(d_total));
    SMVP_CUDA(cudaGetLastError());
    return SMVP_OK;
}

static int tjds_prepare_det(smvp_tjds *A, cudaStream_t s)
{
    if (A->row_exp)
        return SMVP_OK;
    uint32_t *cnt = nullptr;
    SMVP_CUDA(dev_alloc(&A->row_exp, A->rows));
    SMVP_CUDA(dev_alloc(&A->acc, 2 * (int64_t)A->rows));
    SMVP_CUDA(dev_alloc(&cnt, A->rows));
    SMVP_CUDA(cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * (size_t)A->rows, s));
    // EXP_NONE = 0x80808080 is a byte pattern, so a memset fills it
    SMVP_CUDA(cudaMemsetAsync(A->row_exp, 0x80, sizeof(int32_t) * (size_t)A->rows, s));
    if (A->nnz > 0)
    {
        int64_t blocks = ceil_div64(A->nnz, 256 * 4);
        const int64_t cap = (int64_t)device_props().sms * 16;
        if (blocks > cap)
            blocks = cap;
        SMVP_LAUNCH(tjds_row_bound_kernel, (unsigned)blocks, 256, 0, s, mult_rows(A), (const double *)A->val, A->nnz,
                    A->row_exp, cnt); // after the relabel decision: row_exp lives in the index space the kernels use
    }
    if (A->rows > 0)
        SMVP_LAUNCH(tjds_row_exp_kernel, (unsigned)ceil_div64(A->rows, 256), 256, 0, s, A->row_exp, (const uint32_t *)cnt, A->rows);
    SMVP_CUDA(cudaStreamSynchronize(s));
    SMVP_CUDA(cudaFree(cnt));
    SMVP_CUDA(cudaGetLastError());
    return SMVP_OK;
}

} // namespace smvp

using namespace smvp;

extern "C" int smvp_tjds_set_x_device(smvp_tjds *A, const double *d_x, void *stream)
{
    if (!A || (A->cols > 0 && !d_x))
        return SMVP_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (!A->x_exp)
        SMVP_CUDA(dev_alloc(&A->x_exp, 1));
    SMVP_CUDA(cudaMemsetAsync(A->x_exp, 0x80, sizeof(int32_t), s)); // 0x80808080: below every real exponent
    if (A->cols > 0)
    {
        int64_t blocks = ceil_div64(A->cols, 256);
        const int64_t cap = (int64_t)device_props().sms * 8;
        if (blocks > cap)
            blocks = cap;
        SMVP_LAUNCH(tjds_permute_x_kernel, (unsigned)blocks, 256, 0, s, d_x, (const int32_t *)A->perm, A->cols, A->x_perm, A->x_exp);
    }
    SMVP_CUDA(cudaGetLastError());
    return SMVP_OK;
}

// effective number of diagonals to walk: all, or the reference's shipped truncation (see smvp_cuda.h)
static int32_t tjds_effective_limit(const smvp_tjds *A, int32_t diag_limit)
{
    if (diag_limit <= 0)
        return A->ndiag;
    int32_t lim = diag_limit < A->ndiag ? diag_limit : A->ndiag;
    if (lim == A->ndiag && A->ndiag > 0 && A->last_diag_len == 1)
        lim = A->ndiag - 1; // the shipped loop never sees a singleton last diagonal (main-cli.c:957-966)
    return lim;
}

extern "C" int smvp_tjds_mult_device(smvp_tjds *A, double *d_y, int variant, int32_t diag_limit, void *stream)
{
    if (!A || (A->rows > 0 && !d_y))
        return SMVP_E_ARG;
    if (variant != SMVP_TJDS_ATOMIC && variant != SMVP_TJDS_DETERMINISTIC)
        return SMVP_E_ARG;
    if (A->rows == 0)
        return SMVP_OK;
    if (!A->x_exp)
        return SMVP_E_ARG; // smvp_tjds_set_x_device has not been called
    cudaStream_t s = (cudaStream_t)stream;
    const int32_t lim = tjds_effective_limit(A, diag_limit);
    SMVP_TRY(tjds_plan(A, s));
    SMVP_TRY(tjds_relabel_plan(A, s));
    const bool ranked = A->relabel_state == 1;
    const unsigned blocks = (unsigned)A->num_seg_blocks;
    if (variant == SMVP_TJDS_ATOMIC)
    {
        double *y_caller = d_y;
        if (ranked)
            d_y = A->y_rel; // sums land in rank order; tjds_unrank_y_kernel puts them back
        SMVP_CUDA(cudaMemsetAsync(d_y, 0, sizeof(double) * (size_t)A->rows, s));
        if (blocks > 0 && lim > 0)
        {
            const char *ue = getenv("SMVP_TJDS_UNROLL"); // tuning hook; 4 is the measured default
            const int unroll = ue && ue[0] ? atoi(ue) : 4;
            if (unroll == 8)
                SMVP_LAUNCH(tjds_atomic_kernel<8>, blocks, 256, 0, s, (const int2 *)A->seg_blocks, (const int32_t *)A->start_pos, (const int32_t *)A->slot_len,
                        mult_rows(A), (const double *)A->val, (const double *)A->x_perm, d_y, A->nslots, lim);
            else if (unroll == 2)
                SMVP_LAUNCH(tjds_atomic_kernel<2>, blocks, 256, 0, s, (const int2 *)A->seg_blocks, (const int32_t *)A->start_pos, (const int32_t *)A->slot_len,
                        mult_rows(A), (const double *)A->val, (const double *)A->x_perm, d_y, A->nslots, lim);
            else
                SMVP_LAUNCH(tjds_atomic_kernel<4>, blocks, 256, 0, s, (const int2 *)A->seg_blocks, (const int32_t *)A->start_pos, (const int32_t *)A->slot_len,
                        mult_rows(A), (const double *)A->val, (const double *)A->x_perm, d_y, A->nslots, lim);
        }
        if (ranked)
        {
            int64_t ub = ceil_div64(A->rows, 256 * 4);
            const int64_t cap = (int64_t)device_props().sms * 8;
            SMVP_LAUNCH(tjds_unrank_y_kernel, (unsigned)(ub < cap ? ub : cap), 256, 0, s, (const double *)A->y_rel,
                        (const int32_t *)A->row_rank, A->rows, y_caller);
        }
    }
    else
    {
        SMVP_TRY(tjds_prepare_det(A, s));
        SMVP_CUDA(cudaMemsetAsync(A->acc, 0, sizeof(long long) * 2 * (size_t)A->rows, s));
        if (blocks > 0 && lim > 0)
            SMVP_LAUNCH(tjds_det_kernel<4>, blocks, 256, 0, s, (const int2 *)A->seg_blocks, (const int32_t *)A->start_pos, (const int32_t *)A->slot_len,
                        mult_rows(A), (const double *)A->val, (const double *)A->x_perm, (const int32_t *)A->row_exp,
                        (const int32_t *)A->x_exp, (unsigned long long *)A->acc, A->rows, A->nslots, lim);
        SMVP_LAUNCH(tjds_det_finalize_kernel, (unsigned)ceil_div64(A->rows, 256), 256, 0, s, (const long long *)A->acc,
                    (const int32_t *)A->row_exp, (const int32_t *)A->x_exp, A->rows, d_y,
                    ranked ? (const int32_t *)A->row_rank : nullptr);
    }
    SMVP_CUDA(cudaGetLastError());
    return SMVP_OK;
}

extern "C" int smvp_tjds_mult(smvp_tjds *A, const double *x_host, double *y_host, int iters, double *ms_each, int variant,
                              int32_t diag_limit)
{
    if (!A || iters < 1 || (A->cols > 0 && !x_host) || (A->rows > 0 && !y_host))
        return SMVP_E_ARG;
    if (variant != SMVP_TJDS_ATOMIC && variant != SMVP_TJDS_DETERMINISTIC)
        return SMVP_E_ARG;
    if (!A->d_x)
        SMVP_CUDA(dev_alloc(&A->d_x, A->cols));
    if (!A->d_y)
        SMVP_CUDA(dev_alloc(&A->d_y, A->rows));
    if (A->cols > 0)
        SMVP_CUDA(cudaMemcpy(A->d_x, x_host, sizeof(double) * (size_t)A->cols, cudaMemcpyHostToDevice));
    // x is permuted once, before the loop, as the reference does at build time (main-cli.c:907-923)
    SMVP_TRY(smvp_tjds_set_x_device(A, A->d_x, nullptr));
    SMVP_TRY(tjds_plan(A, 0));
    if (A->rows > 0)
        SMVP_TRY(tjds_relabel_plan(A, 0));
    if (variant == SMVP_TJDS_DETERMINISTIC && A->rows > 0)
        SMVP_TRY(tjds_prepare_det(A, 0));
    cudaEvent_t e0, e1;
    SMVP_CUDA(cudaEventCreate(&e0));
    SMVP_CUDA(cudaEventCreate(&e1));
    int rc = SMVP_OK;
    for (int it = 0; it < iters && rc == SMVP_OK; it++)
    {
        // the zero-fill of y / of the accumulators is part of smvp_tjds_mult_device and therefore INSIDE
        // this bracket (the reference keeps its vectorInit outside, main-cli.c:1008): conservative.
        cudaEventRecord(e0, 0);
        rc = smvp_tjds_mult_device(A, A->d_y, variant, diag_limit, nullptr);
        cudaEventRecord(e1, 0);
        if (rc != SMVP_OK)
            break;
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess)
        {
            rc = cuda_fail(e, "cudaEventSynchronize", __FILE__, __LINE__);
            break;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms_each)
            ms_each[it] = (double)ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc != SMVP_OK)
        return rc;
    if (A->rows > 0)
        SMVP_CUDA(cudaMemcpy(y_host, A->d_y, sizeof(double) * (size_t)A->rows, cudaMemcpyDeviceToHost));
    return SMVP_OK;
}

extern "C" int smvp_tjds_info(const smvp_tjds *A, smvp_tjds_info_t *out)
{
    if (!A || !out)
        return SMVP_E_ARG;
    out->rows = A->rows;
    out->cols = A->cols;
    out->nnz = A->nnz;
    out->ndiag = A->ndiag;
    out->ref_diag_limit = A->ref_diag_limit;
    out->input_order = A->input_order;
    out->bytes_per_mult = 12 * A->nnz + 4 * ((int64_t)A->ndiag + 1) + 8 * (int64_t)A->cols + 8 * (int64_t)A->rows;
    out->device_bytes = A->device_bytes;
    out->launches_per_mult[SMVP_TJDS_ATOMIC] = A->relabel_state == 1 ? 2 : 1;
    out->launches_per_mult[SMVP_TJDS_DETERMINISTIC] = 2;
    out->y_relabel = A->relabel_state;
    return SMVP_OK;
}
