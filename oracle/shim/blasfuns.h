/* TEST INFRASTRUCTURE ONLY (oracle build shim) - not part of the product.
 *
 * The reference core (/root/reference/src/stochqn.c:79) does `#include "blasfuns.h"`
 * when it is built outside R / Python.  Upstream generates that header with CMake
 * (src/blasfuns.h.in); we do not run the reference's build system, so this file
 * stands in for it and routes the ten CBLAS entry points the core uses
 * (stochqn.c:289,676-678,686-688,698,705-706,829,838,892,923,946-949,1006) onto the
 * LP64 OpenBLAS that ships inside the SciPy wheel (symbols are prefixed `scipy_`).
 * `real_t` and the cblas_t* aliases come from the reference's own header
 * (include/stochqn.h:62-76), which the core includes before this file.
 */
#ifndef STOCHQN_ORACLE_BLAS_SHIM_H
#define STOCHQN_ORACLE_BLAS_SHIM_H

typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE {
    CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113, CblasConjNoTrans = 114
} CBLAS_TRANSPOSE;

#define cblas_ddot  scipy_cblas_ddot
#define cblas_daxpy scipy_cblas_daxpy
#define cblas_dscal scipy_cblas_dscal
#define cblas_dnrm2 scipy_cblas_dnrm2
#define cblas_dgemv scipy_cblas_dgemv
#define cblas_sdot  scipy_cblas_sdot
#define cblas_saxpy scipy_cblas_saxpy
#define cblas_sscal scipy_cblas_sscal
#define cblas_snrm2 scipy_cblas_snrm2
#define cblas_sgemv scipy_cblas_sgemv

real_t cblas_tdot(const int n, const real_t *x, const int incx, const real_t *y, const int incy);
void   cblas_taxpy(const int n, const real_t alpha, const real_t *x, const int incx, real_t *y, const int incy);
void   cblas_tscal(const int n, const real_t alpha, real_t *x, const int incx);
real_t cblas_tnrm2(const int n, const real_t *x, const int incx);
void   cblas_tgemv(const CBLAS_ORDER order, const CBLAS_TRANSPOSE trans, const int m, const int n,
                   const real_t alpha, const real_t *a, const int lda, const real_t *x, const int incx,
                   const real_t beta, real_t *y, const int incy);

#endif
