/* stand-in for <R_ext/BLAS.h> (oracle test infrastructure only).
 * Only the two level-1 routines the reference calls (src/matmul.cpp:45,50,72) are declared; they are
 * defined as plain loops in oracle/ref_driver.cpp (reference-BLAS semantics, no FMA contraction
 * unless the compiler flags enable it). saxpy/scopy are defined by src/matmul.cpp itself (17-40). */
#ifndef F77_NAME
#define F77_CALL(x) x##_
#define F77_NAME(x) x##_
#endif
void F77_NAME(daxpy)(const int *n, const double *da, const double *dx, const int *incx,
                     double *dy, const int *incy);
void F77_NAME(dcopy)(const int *n, const double *dx, const int *incx, double *dy, const int *incy);
