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
 * For src/misc.cpp and src/operators.cpp (sort / validity checks / elementwise products) the
 * subset is a little wider: iterator-range constructors, Rcpp::String, Rcpp::stop, LogicalMatrix, List["name"]
 * assignment and lookup, ISNAN/ISNA/R_FINITE/R_NaN/R_pow (R_pow is std::pow here: only there so that the whole file
 * links; no checked function of the oracle goes through it).
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
#include <cfloat>
#include <limits>
#include <stdexcept>
#include <iterator>
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
static inline bool mx_shim_isna(double x)
{
    uint64_t bits;
    std::memcpy(&bits, &x, sizeof(bits));
    return std::isnan(x) && (uint32_t)bits == 1954u; /* arithmetic.c: R_IsNA */
}
#ifndef ISNAN
#define ISNAN(x) (std::isnan((double)(x)))
#endif
#ifndef ISNA
#define ISNA(x) (mx_shim_isna((double)(x)))
#endif
#ifndef R_FINITE
#define R_FINITE(x) (std::isfinite((double)(x)))
#endif
#ifndef R_NaN
#define R_NaN (std::numeric_limits<double>::quiet_NaN())
#endif
#ifndef R_PosInf
#define R_PosInf (std::numeric_limits<double>::infinity())
#define R_NegInf (-std::numeric_limits<double>::infinity())
#endif
static inline double R_pow(double x, double y) { return std::pow(x, y); }

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
    /* storage from a custom allocator; the deleter runs when the last handle goes (R: Rf_allocVector3 + garbage collection) */
    ShimVector(std::shared_ptr<T> owned, size_t n) : own_(owned), ptr_(owned.get()), n_(n) {}
    /* Rcpp's copying iterator-range constructor */
    template <class It, class = typename std::enable_if<!std::is_integral<It>::value>::type,
              class = typename std::iterator_traits<It>::value_type>
    ShimVector(It first, It last) : n_((size_t)std::distance(first, last))
    {
        own_ = std::shared_ptr<T>(new T[n_ ? n_ : 1](), std::default_delete<T[]>());
        ptr_ = own_.get();
        size_t k = 0;
        for (It it = first; it != last; ++it) ptr_[k++] = (T)*it;
    }
    /* Rcpp vectors convert to SEXP implicitly */
    operator SEXP() const
    {
        SEXP s = new mx_shim_sexprec();
        if (std::is_same<T, double>::value) s->dbls.assign(ptr_, ptr_ + n_);
        else { s->is_int = true; s->ints.assign(ptr_, ptr_ + n_); }
        return s;
    }
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
    ShimMatrix(std::shared_ptr<T> owned, int nrow, int ncol)
        : ShimVector<T, Tag>(owned, (size_t)nrow * (size_t)ncol), nrow_(nrow), ncol_(ncol) {}
    int nrow() const { return nrow_; }
    int ncol() const { return ncol_; }
private:
    int nrow_, ncol_;
};

typedef ShimMatrix<double, mx_tag_num> NumericMatrix;
typedef ShimMatrix<int, mx_tag_int> IntegerMatrix;
typedef ShimMatrix<int, mx_tag_lgl> LogicalMatrix;

struct String {
    std::string s;
    String() {}
    String(const char *c) : s(c) {}
    String(const std::string &c) : s(c) {}
};

template <class... Args>
[[noreturn]] static inline void stop(const char *msg, Args...) { throw std::runtime_error(msg); }
static inline void checkUserInterrupt() {}

/* ---- List::create(_["name"] = value, ...) ---- */
struct ListEntry {
    std::string name;
    std::shared_ptr<void> keep;
    void *data = nullptr;
    size_t size = 0;
    bool is_int = false;
    std::string str; /* for Rcpp::String values */
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

class List;
struct ListProxy {
    List *list;
    std::string name;
    template <class T, class Tag> ListProxy &operator=(const ShimVector<T, Tag> &v);
    ListProxy &operator=(SEXP s);
    void *data_ptr() const;
};

class List {
public:
    std::vector<ListEntry> entries;
    ListProxy operator[](const char *name) { return ListProxy{this, name}; }
    ListEntry *find(const std::string &name)
    {
        for (auto &e : entries) if (e.name == name) return &e;
        return nullptr;
    }
    template <class T, class Tag>
    void set(const std::string &name, const ShimVector<T, Tag> &v)
    {
        ListEntry *e = find(name);
        if (!e) { entries.push_back(ListEntry()); e = &entries.back(); e->name = name; }
        e->keep = v.keepalive();
        e->data = (void *)v.data_ptr();
        e->size = (size_t)v.size();
        e->is_int = std::is_same<T, int>::value;
    }
    template <class... Args>
    static List create(const Args &...args)
    {
        List out;
        int dummy[] = {0, (out.push(args), 0)...};
        (void)dummy;
        return out;
    }
private:
    template <class V>
    void push(const NamedValue<V> &nv) { store(nv.name, nv.value); }
    template <class T, class Tag>
    void store(const char *name, const ShimVector<T, Tag> &v) { set(name, v); }
    void store(const char *name, const String &v)
    {
        ListEntry e;
        e.name = name;
        e.str = v.s;
        entries.push_back(e);
    }
    template <class V, class = typename std::enable_if<std::is_arithmetic<V>::value>::type>
    void store(const char *name, V)
    {
        ListEntry e;
        e.name = name;
        entries.push_back(e);
    }
};

template <class T, class Tag>
inline ListProxy &ListProxy::operator=(const ShimVector<T, Tag> &v) { list->set(name, v); return *this; }
inline ListProxy &ListProxy::operator=(SEXP s)
{
    if (s->is_int) list->set(name, IntegerVector(s));
    else list->set(name, NumericVector(s));
    return *this;
}
inline void *ListProxy::data_ptr() const
{
    ListEntry *e = list->find(name);
    return e ? e->data : nullptr;
}

template <class Fn>
SEXP unwindProtect(Fn fn, void *arg) { return fn(arg); }

/* External pointers (for the device-resident handles of rglue/handle_gpu_glue.cpp): a reference-counted owner
 * that runs the finalizer when the last copy goes away (R: when the object is garbage-collected) or on release(). */
template <class T> struct PreserveStorage {};
template <class T> void standard_delete_finalizer(T *obj) { delete obj; }
template <class T, template <class> class StoragePolicy = PreserveStorage, void Finalizer(T *) = standard_delete_finalizer<T>,
          bool finalizeOnExit = false>
class XPtr {
public:
    explicit XPtr(T *p, bool set_delete_finalizer = true) : box_(std::make_shared<Box>(p, set_delete_finalizer)) {}
    T *get() const { return box_->ptr; }
    T *checked_get() const
    {
        if (!box_->ptr) throw std::runtime_error("external pointer is not valid");
        return box_->ptr;
    }
    T *operator->() const { return checked_get(); }
    void release() { box_->finalize(); }

private:
    struct Box {
        T *ptr;
        bool fin;
        Box(T *p, bool f) : ptr(p), fin(f) {}
        void finalize()
        {
            if (ptr && fin) Finalizer(ptr);
            ptr = nullptr;
        }
        ~Box() { finalize(); }
    };
    std::shared_ptr<Box> box_;
};

} /* namespace Rcpp */

template <class V> static inline int *INTEGER(const V &v) { return (int *)v.data_ptr(); }
template <class V> static inline int *LOGICAL(const V &v) { return (int *)v.data_ptr(); }
template <class V> static inline double *REAL(const V &v) { return (double *)v.data_ptr(); }

#endif
