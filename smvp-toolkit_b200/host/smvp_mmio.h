/*
 * smvp_mmio.h -- Matrix Market banner / size-line reader with the interface of the NIST "mmio" library the
 * reference links (reference: mmio/mmio.h; used at main-cli.c:1405 mm_read_banner, :1419
 * mm_read_mtx_crd_size and through the mm_is_* predicates at :1410, :1429).
 *
 * Same names, same typecode convention (4 characters: object, format, field, symmetry), same error
 * codes, so host code written against the NIST header compiles against this one.  The implementation
 * (smvp_mmio.c) is written from the Matrix Market exchange-format specification, not from the NIST source.
 */
#ifndef SMVP_MMIO_H
#define SMVP_MMIO_H

#include <stdio.h>

#define MM_MAX_LINE_LENGTH 1025
#define MM_MAX_TOKEN_LENGTH 64
#define MatrixMarketBanner "%%MatrixMarket"

typedef char MM_typecode[4];

/* error codes (values of the NIST library) */
#define MM_COULD_NOT_READ_FILE 11
#define MM_PREMATURE_EOF 12
#define MM_NOT_MTX 13
#define MM_NO_HEADER 14
#define MM_UNSUPPORTED_TYPE 15
#define MM_LINE_TOO_LONG 16
#define MM_COULD_NOT_WRITE_FILE 17

/* typecode[0]: 'M' matrix; [1]: 'C' coordinate / 'A' array; [2]: 'R' real, 'C' complex, 'P' pattern,
 * 'I' integer; [3]: 'G' general, 'S' symmetric, 'K' skew-symmetric, 'H' hermitian */
#define mm_is_matrix(t) ((t)[0] == 'M')
#define mm_is_sparse(t) ((t)[1] == 'C')
#define mm_is_coordinate(t) ((t)[1] == 'C')
#define mm_is_dense(t) ((t)[1] == 'A')
#define mm_is_array(t) ((t)[1] == 'A')
#define mm_is_complex(t) ((t)[2] == 'C')
#define mm_is_real(t) ((t)[2] == 'R')
#define mm_is_pattern(t) ((t)[2] == 'P')
#define mm_is_integer(t) ((t)[2] == 'I')
#define mm_is_symmetric(t) ((t)[3] == 'S')
#define mm_is_general(t) ((t)[3] == 'G')
#define mm_is_skew(t) ((t)[3] == 'K')
#define mm_is_hermitian(t) ((t)[3] == 'H')

#define mm_set_matrix(t) ((*(t))[0] = 'M')
#define mm_set_coordinate(t) ((*(t))[1] = 'C')
#define mm_set_sparse(t) mm_set_coordinate(t)
#define mm_set_array(t) ((*(t))[1] = 'A')
#define mm_set_dense(t) mm_set_array(t)
#define mm_set_complex(t) ((*(t))[2] = 'C')
#define mm_set_real(t) ((*(t))[2] = 'R')
#define mm_set_pattern(t) ((*(t))[2] = 'P')
#define mm_set_integer(t) ((*(t))[2] = 'I')
#define mm_set_symmetric(t) ((*(t))[3] = 'S')
#define mm_set_general(t) ((*(t))[3] = 'G')
#define mm_set_skew(t) ((*(t))[3] = 'K')
#define mm_set_hermitian(t) ((*(t))[3] = 'H')
#define mm_clear_typecode(t) ((*(t))[0] = (*(t))[1] = (*(t))[2] = ' ', (*(t))[3] = 'G')
#define mm_initialize_typecode(t) mm_clear_typecode(t)

int mm_read_banner(FILE *f, MM_typecode *matcode);
int mm_read_mtx_crd_size(FILE *f, int *M, int *N, int *nz);
int mm_write_banner(FILE *f, MM_typecode matcode);
int mm_write_mtx_crd_size(FILE *f, int M, int N, int nz);
int mm_is_valid(MM_typecode matcode);
/* returns a pointer to a static buffer ("matrix coordinate real general") */
char *mm_typecode_to_str(MM_typecode matcode);

#endif
