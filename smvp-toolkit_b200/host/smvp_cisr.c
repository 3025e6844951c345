/*
 * smvp_cisr.c -- CISR (condensed interleaved sparse representation) .coe emitter, the reference's `-g`
 * option (smvp_cisr_coegen, main-cli.c:473-729).  Host-only integer packing for a Xilinx BRAM image; it
 * is not on the GPU path and exists so that the command line keeps its whole surface.
 *
 * Restated from the reference's behaviour:
 *   scheduling (:540-612)  `slots` channels each stream one CSR row at a time; slot group g holds, per
 *       channel, the index of the nonzero it emits in cycle g.  A channel whose previous index was the
 *       last of its row (index >= row_end - 1) takes the next unassigned row, or the invalid index nnz + 1
 *       when no rows are left.  Groups are produced until every channel is invalid.  Row lengths are
 *       recorded in the order rows are handed out (= row order).
 *   expansion (:628-654)   invalid entries are padded with value 0, column 0.
 *   packing (:690-728)     36-bit words printed as 2 + 8 hex digits:
 *       00 AAAAAAAA                      start of data
 *       01 (int)val << 20 | col << 8 | slot          one per scheduled entry
 *       02 1<<28 | len_a << 16 | valid_b << 12 | len_b    after an entry word while row lengths remain
 *       03 FFFFFFFF                      end of data
 */
#include "smvp_host.h"

#include <stdio.h>
#include <stdlib.h>

int smvp_cisr_coe(FILE *out, const int32_t *row_ptr, const int32_t *col_ind, const double *val, int rows, int64_t nnz_,
                  int slots)
{
    const int nnz = (int)nnz_;
    const int invalid = nnz + 1; /* main-cli.c:561: row_ptr[rows] + 1 */
    int *cur, *row_end, *sched = NULL;
    size_t cap = 0, ngroups = 0;
    int next_row = 0, all_invalid = 0, s;
    size_t g, k;
    int rl = 0;

    if (!out || slots < 1 || rows < 0 || nnz < 0 || (rows > 0 && !row_ptr))
        return -1;
    cur = (int *)malloc(sizeof(int) * (size_t)slots);
    row_end = (int *)calloc((size_t)slots, sizeof(int));
    if (!cur || !row_end)
    {
        free(cur);
        free(row_end);
        return -1;
    }

    while (!all_invalid)
    {
        if ((ngroups + 1) * (size_t)slots > cap)
        {
            int *ns;
            cap = cap ? cap * 2 : (size_t)slots * 64;
            ns = (int *)realloc(sched, sizeof(int) * cap);
            if (!ns)
            {
                free(sched);
                free(cur);
                free(row_end);
                return -1;
            }
            sched = ns;
        }
        for (s = 0; s < slots; s++)
        {
            /* first group: every channel takes a fresh row; later: only channels that finished theirs */
            const int take_row = (ngroups == 0) || (cur[s] >= row_end[s] - 1);
            if (!take_row)
                cur[s] = cur[s] + 1;
            else if (next_row < rows)
            {
                cur[s] = row_ptr[next_row];
                row_end[s] = row_ptr[next_row + 1];
                next_row++;
            }
            else
                cur[s] = invalid;
            sched[ngroups * (size_t)slots + (size_t)s] = cur[s];
        }
        all_invalid = 1;
        for (s = 0; s < slots; s++)
            if (cur[s] < nnz)
                all_invalid = 0;
        ngroups++;
        if (ngroups >= (size_t)(nnz > 0 ? nnz : 1) && !all_invalid)
        {
            /* main-cli.c:607-611: the reference aborts here */
            free(sched);
            free(cur);
            free(row_end);
            return -2;
        }
    }

    fprintf(out, "\n;*********************************************");
    fprintf(out, "\n;* CISR COE File for Vivado Single-Port BRAM *");
    fprintf(out, "\n;*********************************************\n");
    fprintf(out, "\n;Generated with a slot/channel count of: %d\n\n", slots);
    fprintf(out, "memory_initialization_radix=16;\n");
    fprintf(out, "memory_initialization_vector=\n");
    fprintf(out, "00%08x,\n", 0xAAAAAAAAu);
    for (g = 0; g < ngroups; g++)
        for (k = 0; k < (size_t)slots; k++)
        {
            const int idx = sched[g * (size_t)slots + k];
            const int v = idx >= nnz ? 0 : (int)val[idx];
            const int c = idx >= nnz ? 0 : col_ind[idx];
            unsigned word = ((unsigned)v << 20) | ((unsigned)c << 8) | (unsigned)k;
            fprintf(out, "01%08x,\n", word);
            if (rl < rows)
            {
                word = (1u << 28) | ((unsigned)(row_ptr[rl + 1] - row_ptr[rl]) << 16);
                rl++;
                if (rl < rows)
                {
                    word |= (1u << 12) | (unsigned)(row_ptr[rl + 1] - row_ptr[rl]);
                    rl++;
                }
                fprintf(out, "02%08x,\n", word);
            }
        }
    fprintf(out, "03%08x;\n\n", 0xFFFFFFFFu);
    free(sched);
    free(cur);
    free(row_end);
    return 0;
}
