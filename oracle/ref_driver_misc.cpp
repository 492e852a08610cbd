/* oracle/ref_driver_misc.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Compiles the REFERENCE's src/misc.cpp in place (found through -I/root/reference/src, nothing copied)
 * into oracle/_ref/libmxref_ops.so and exposes, with plain pointers for ctypes, the functions SURVEY.md
 * §8 f3 names: per-row index sorting and the CSR validity checks (src/misc.cpp:117-127, 161-228,
 * 300-330, 970-1016; R callers R/utils.R:22-161, 439-489).
 */
#include "misc.cpp"

#include <cstdio>

typedef Rcpp::IntegerVector IV;
typedef Rcpp::NumericVector NV;

extern "C" {

/* src/misc.cpp:300-313 sort_sparse_indices_numeric: sorts the caller's arrays IN PLACE (rows already sorted are
 * left alone, 213-214). */
int mxref_sort_sparse_indices_numeric(const int *p, int nrows, int *j, double *x, int nnz)
{
    sort_sparse_indices_numeric(IV((int *)p, (size_t)nrows + 1), IV(j, (size_t)nnz), NV(x, (size_t)nnz));
    return 0;
}

/* src/misc.cpp:161-175 (int* overload): true when EVERY row is sorted, despite the name. */
int mxref_check_indices_are_unsorted(const int *p, int nrows, const int *j)
{
    return check_indices_are_unsorted((int *)p, (int *)j, nrows) ? 1 : 0;
}

/* src/misc.cpp:970-1016: returns 0 when the list is empty, else 1 and the "err" string. */
int mxref_check_valid_csr_matrix(const int *p, int nrows, const int *j, int nnz, int ncols, char *err, int errlen)
{
    Rcpp::List res = check_valid_csr_matrix(IV((int *)p, (size_t)nrows + 1), IV((int *)j, (size_t)nnz), nrows, ncols);
    if (res.entries.empty()) return 0;
    if (err && errlen > 0) std::snprintf(err, (size_t)errlen, "%s", res.entries[0].str.c_str());
    return 1;
}

} /* extern "C" */
