/* Stand-in for <Rcpp.h>: TEST INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
 *
 * Purpose: let the reference's own src/matmul.cpp (and src/MatrixExtra.h, which it includes by
 * quote-include from its own directory) compile UNMODIFIED and IN PLACE from /root/reference/src,
 * without R or Rcpp being installed.  Only the tiny subset of the Rcpp API that file touches is
 * provided: Integer/Numeric/LogicalVector, Numeric/IntegerMatrix, List::create(_["x"]=...),
 * unwindProtect, INTEGER()/REAL()/LOGICAL(), NA_* constants.
 *
 * Semantics preserved from real Rcpp that the reference relies on:
 *   - Vector(n) / Matrix(nr, nc) allocate ZERO-FILLED storage (src/matmul.cpp:197,261,323,389,494);
 *   - copies share storage (Rcpp vectors are handles on a SEXP);
 *   - IntegerVector and LogicalVector are DISTINCT types (std::is_same dispatch at
 *     src/matmul.cpp:406-411 depends on it);
 *   - matrices are column-major.
 * Added for the ctypes driver: non-owning (pointer, size) constructors that borrow caller memory,
 * exactly like Rcpp borrows the SEXP payload.
 */
#ifndef MX_ORACLE_SHIM_RCPP_H
#define MX_ORACLE_SHIM_RCPP_H

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <climits>
#include <cmath>
#include <memory>
#include <vector>
#include <string>
#include <utility>
#include <type_traits>

/* ---- minimal "R internals" ---- */
struct mx_shim_sexprec {
    std::vector<int> ints;
    std::vector<double> dbls;
    bool is_int = false;
};
typedef mx_shim_sexprec *SEXP;

#ifndef NA_INTEGER
#define NA_INTEGER INT_MIN
#endif
#ifndef NA_LOGICAL
#define NA_LOGICAL INT_MIN
#endif
static inline double mx_shim_na_real()
{
    /* R's NA_real_: quiet NaN whose low word is 1954 (arithmetic.c: R_ValueOfNA) */
    const uint64_t bits = 0x7FF00000000007A2ULL;
    double out;
    std::memcpy(&out, &bits, sizeof(out));
    return out;
}
#ifndef NA_REAL
#define NA_REAL (mx_shim_na_real())
#endif

namespace Rcpp {

struct mx_tag_int {};
struct mx_tag_lgl {};
struct mx_tag_num {};

template <class T, class Tag>
class ShimVector {
public:
    typedef T stored_type;
    ShimVector() : ptr_(nullptr), n_(0) {}
    template <class I, class = typename std::enable_if<std::is_integral<I>::value>::type>
    explicit ShimVector(I n) : n_((size_t)n)
    {
        own_ = std::shared_ptr<T>(new T[n_ ? n_ : 1](), std::default_delete<T[]>());
        ptr_ = own_.get();
    }
    /* non-owning view over caller memory (what Rcpp does with a SEXP) */
    ShimVector(T *borrowed, size_t n) : ptr_(borrowed), n_(n) {}
    /* from the fake SEXP returned by SafeRcppVector / unwindProtect */
    ShimVector(SEXP s)
    {
        if (std::is_same<T, double>::value) {
            n_ = s->dbls.size();
            own_ = std::shared_ptr<T>(new T[n_ ? n_ : 1](), std::default_delete<T[]>());
            for (size_t i = 0; i < n_; i++) own_.get()[i] = (T)s->dbls[i];
        } else {
            n_ = s->ints.size();
            own_ = std::shared_ptr<T>(new T[n_ ? n_ : 1](), std::default_delete<T[]>());
            for (size_t i = 0; i < n_; i++) own_.get()[i] = (T)s->ints[i];
        }
        ptr_ = own_.get();
        delete s;
    }
    long size() const { return (long)n_; }
    long length() const { return (long)n_; }
    template <class I> T &operator[](I i) { return ptr_[(size_t)i]; }
    template <class I> const T &operator[](I i) const { return ptr_[(size_t)i]; }
    T *begin() { return ptr_; }
    T *end() { return ptr_ + n_; }
    T *data_ptr() const { return ptr_; }
    std::shared_ptr<T> keepalive() const { return own_; }
protected:
    std::shared_ptr<T> own_;
    T *ptr_;
    size_t n_;
};

typedef ShimVector<int, mx_tag_int> IntegerVector;
typedef ShimVector<int, mx_tag_lgl> LogicalVector;
typedef ShimVector<double, mx_tag_num> NumericVector;

template <class T, class Tag>
class ShimMatrix : public ShimVector<T, Tag> {
public:
    ShimMatrix() : nrow_(0), ncol_(0) {}
    ShimMatrix(int nrow, int ncol)
        : ShimVector<T, Tag>((size_t)nrow * (size_t)ncol), nrow_(nrow), ncol_(ncol) {}
    ShimMatrix(T *borrowed, int nrow, int ncol)
        : ShimVector<T, Tag>(borrowed, (size_t)nrow * (size_t)ncol), nrow_(nrow), ncol_(ncol) {}
    int nrow() const { return nrow_; }
    int ncol() const { return ncol_; }
private:
    int nrow_, ncol_;
};

typedef ShimMatrix<double, mx_tag_num> NumericMatrix;
typedef ShimMatrix<int, mx_tag_int> IntegerMatrix;

/* ---- List::create(_["name"] = value, ...) ---- */
struct ListEntry {
    std::string name;
    std::shared_ptr<void> keep;
    void *data = nullptr;
    size_t size = 0;
    bool is_int = false;
};

template <class V>
struct NamedValue {
    const char *name;
    V value;
};

struct NamedPlaceholderItem {
    const char *name;
    template <class V>
    NamedValue<V> operator=(const V &v) const { return NamedValue<V>{name, v}; }
};

struct NamedPlaceholder {
    NamedPlaceholderItem operator[](const char *name) const { return NamedPlaceholderItem{name}; }
};
static const NamedPlaceholder _ = NamedPlaceholder();

class List {
public:
    std::vector<ListEntry> entries;
    template <class... Args>
    static List create(const Args &...args)
    {
        List out;
        int dummy[] = {0, (out.push(args), 0)...};
        (void)dummy;
        return out;
    }
private:
    template <class T, class Tag>
    void push(const NamedValue<ShimVector<T, Tag>> &nv)
    {
        ListEntry e;
        e.name = nv.name;
        e.keep = nv.value.keepalive();
        e.data = (void *)nv.value.data_ptr();
        e.size = (size_t)nv.value.size();
        e.is_int = std::is_same<T, int>::value;
        entries.push_back(e);
    }
};

template <class Fn>
SEXP unwindProtect(Fn fn, void *arg) { return fn(arg); }

} /* namespace Rcpp */

template <class V> static inline int *INTEGER(const V &v) { return (int *)v.data_ptr(); }
template <class V> static inline int *LOGICAL(const V &v) { return (int *)v.data_ptr(); }
template <class V> static inline double *REAL(const V &v) { return (double *)v.data_ptr(); }

#endif
