/* rglue/mxgpu_result_alloc.h — how the glue allocates the matrices and vectors it RETURNS.
 *
 * The reference's exports return a freshly allocated, zero-filled R matrix (src/matmul.cpp:197, 261, 323, 389).  On the GPU
 * path every element is written by the device-to-host copy, so nothing has to be zero-filled — but a freshly allocated
 * R matrix still consists of pages that do not exist yet, and copying 512 MB into them costs 27 ms on the B200 host
 * (kernel page zeroing + a bounce through a page-locked slot), half of a cfg3 call.  R lets a package supply the
 * allocator of a vector: Rf_allocVector3(type, n, R_allocator_t*).  Large results are therefore allocated from the
 * library's pool of PAGE-LOCKED blocks (mxg_host_alloc): the device writes the result straight into the R object, and
 * when R's garbage collector frees the matrix the allocator's free hook hands the block back to the pool
 * (mxg_host_free), where the next result of that size reuses it.  To R it is an ordinary matrix.
 * When the pool is full or unavailable the hook falls back to malloc (and the library to its bounce path).
 *
 * Test builds (-DMXGPU_GLUE_SHIM, no R): the same pool through the stand-in's shared-ownership constructor, whose
 * deleter plays the garbage collector.
 */
#ifndef MXGPU_RESULT_ALLOC_H
#define MXGPU_RESULT_ALLOC_H

#include <cstddef>
#include <cstdlib>

#include "mxgpu.h"

#ifndef MXGPU_PINNED_RESULT_MIN
#define MXGPU_PINNED_RESULT_MIN ((size_t)1 << 20) /* smaller results stay on R's own heap */
#endif

#if defined(MXGPU_GLUE_SHIM)

#include <memory>

template <class Matrix>
inline Matrix mxgpu_new_matrix(int nrow, int ncol)
{
    typedef typename Matrix::stored_type T;
    const size_t n = (size_t)nrow * (size_t)ncol;
    void *p = NULL;
    if (n * sizeof(T) >= MXGPU_PINNED_RESULT_MIN && mxg_host_alloc(n * sizeof(T), &p) == MXG_OK && p)
        return Matrix(std::shared_ptr<T>(static_cast<T *>(p), [](T *q) { mxg_host_free(q); }), nrow, ncol);
    return Matrix(nrow, ncol);
}

template <class Vector>
inline Vector mxgpu_new_vector(size_t n)
{
    typedef typename Vector::stored_type T;
    void *p = NULL;
    if (n * sizeof(T) >= MXGPU_PINNED_RESULT_MIN && mxg_host_alloc(n * sizeof(T), &p) == MXG_OK && p)
        return Vector(std::shared_ptr<T>(static_cast<T *>(p), [](T *q) { mxg_host_free(q); }), n);
    return Vector(n);
}

#else /* the real thing */

#include <Rcpp.h>
#include <R_ext/Rallocators.h>

namespace mxgpu_alloc {

inline void *hook_alloc(R_allocator_t *, size_t bytes)
{
    void *p = NULL;
    if (mxg_host_alloc(bytes, &p) == MXG_OK && p) return p;
    return std::malloc(bytes); /* pool full / no device: an ordinary block, recognised again in hook_free */
}

inline void hook_free(R_allocator_t *, void *p)
{
    if (mxg_host_free(p) != MXG_OK) std::free(p);
}

/* R copies the allocator struct next to the vector it allocates, so a static instance is enough */
inline SEXP big_vector(SEXPTYPE type, R_xlen_t n)
{
    static R_allocator_t hooks = {hook_alloc, hook_free, NULL, NULL};
    return Rf_allocVector3(type, n, &hooks);
}

} /* namespace mxgpu_alloc */

/* no zero fill anywhere: every element is written by the device-to-host copy */
template <class Matrix>
inline Matrix mxgpu_new_matrix(int nrow, int ncol)
{
    typedef typename Matrix::stored_type T;
    const R_xlen_t n = (R_xlen_t)nrow * (R_xlen_t)ncol;
    if ((size_t)n * sizeof(T) < MXGPU_PINNED_RESULT_MIN) return Matrix(Rcpp::no_init(nrow, ncol));
    Rcpp::Shield<SEXP> v(mxgpu_alloc::big_vector(Rcpp::traits::r_sexptype_traits<T>::rtype, n));
    Rcpp::Shield<SEXP> dim(Rf_allocVector(INTSXP, 2));
    INTEGER(dim)[0] = nrow;
    INTEGER(dim)[1] = ncol;
    Rf_setAttrib(v, R_DimSymbol, dim);
    return Matrix(static_cast<SEXP>(v));
}

template <class Vector>
inline Vector mxgpu_new_vector(size_t n)
{
    typedef typename Vector::stored_type T;
    if (n * sizeof(T) < MXGPU_PINNED_RESULT_MIN) return Vector(Rcpp::no_init((R_xlen_t)n));
    Rcpp::Shield<SEXP> v(mxgpu_alloc::big_vector(Rcpp::traits::r_sexptype_traits<T>::rtype, (R_xlen_t)n));
    return Vector(static_cast<SEXP>(v));
}

#endif
#endif /* MXGPU_RESULT_ALLOC_H */
