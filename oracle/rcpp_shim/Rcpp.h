// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Stand-in for <Rcpp.h>: just enough of the Rcpp surface for the reference translation unit
// (/root/reference/src/microclimfCpp.cpp + microclimfheaders.h) to compile UNCHANGED with g++ and run
// outside R.  It lets oracle/Makefile build oracle/_ref/libmicroclimf_ref.so straight from the
// reference sources where they lie (nothing from the reference is copied into this repository).
//
// Semantics that matter for fidelity:
//   * vectors/matrices are reference-counted handles: copying a handle aliases the storage;
//   * NA_REAL is R's NA bit pattern (quiet NaN, low word 1954); is_na() is true for ANY NaN,
//     exactly as Rcpp's traits::is_na<REALSXP> (R_isnancpp);
//   * NumericVector(n) zero-fills; NumericVector(n, v) fills with v;
//   * matrices are column-major: (i, j) -> i + nrow * j.
#ifndef MCF_ORACLE_RCPP_SHIM_H
#define MCF_ORACLE_RCPP_SHIM_H

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

typedef std::ptrdiff_t R_xlen_t;

inline double mcf_shim_na_real() {
    const uint64_t bits = 0x7FF00000000007A2ULL;
    double d;
    std::memcpy(&d, &bits, sizeof d);
    return d;
}
static const double NA_REAL = mcf_shim_na_real();
static const int NA_INTEGER = INT_MIN;
static const double R_NaN = std::nan("");
static const double R_PosInf = INFINITY;
static const double R_NegInf = -INFINITY;
inline int R_IsNA(double x) {
    if (!std::isnan(x)) return 0;
    uint64_t b;
    std::memcpy(&b, &x, sizeof b);
    return (b & 0xFFFFFFFFULL) == 1954;
}
#define RcppExport extern "C"

namespace Rcpp {

template <class T> class Vector;
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;

// `x.attr("dim")`: usable on both sides of an assignment.
template <class T> class DimProxy {
    Vector<T>& owner_;
public:
    explicit DimProxy(Vector<T>& o) : owner_(o) {}
    DimProxy& operator=(const IntegerVector& d);
    operator IntegerVector() const;
};

template <class T> class Vector {
protected:
    std::shared_ptr<std::vector<T>> d_;
    std::shared_ptr<std::vector<int>> dim_;
public:
    typedef T value_type;
    Vector() : d_(std::make_shared<std::vector<T>>()), dim_(std::make_shared<std::vector<int>>()) {}
    template <class N, class = typename std::enable_if<std::is_integral<N>::value>::type>
    explicit Vector(N n) : d_(std::make_shared<std::vector<T>>((size_t)n, T(0))), dim_(std::make_shared<std::vector<int>>()) {}
    template <class N, class U, class = typename std::enable_if<std::is_integral<N>::value && std::is_arithmetic<U>::value>::type>
    Vector(N n, U fill) : d_(std::make_shared<std::vector<T>>((size_t)n, (T)fill)), dim_(std::make_shared<std::vector<int>>()) {}
    Vector(std::initializer_list<T> il) : d_(std::make_shared<std::vector<T>>(il)), dim_(std::make_shared<std::vector<int>>()) {}
    Vector(const std::vector<T>& v) : d_(std::make_shared<std::vector<T>>(v)), dim_(std::make_shared<std::vector<int>>()) {}
    // coercion between integer and double vectors copies (as R's coerceVector does)
    template <class U, class = typename std::enable_if<!std::is_same<U, T>::value>::type>
    Vector(const Vector<U>& o) : d_(std::make_shared<std::vector<T>>(o.size())), dim_(std::make_shared<std::vector<int>>(o.dims())) {
        for (size_t i = 0; i < d_->size(); ++i) (*d_)[i] = (T)o[i];
    }
    static Vector create(T a) { return Vector({a}); }
    static Vector create(T a, T b) { return Vector({a, b}); }
    static Vector create(T a, T b, T c) { return Vector({a, b, c}); }
    static Vector create(T a, T b, T c, T e) { return Vector({a, b, c, e}); }

    template <class N> T& operator[](N i) { return (*d_)[(size_t)i]; }
    template <class N> const T& operator[](N i) const { return (*d_)[(size_t)i]; }
    R_xlen_t size() const { return (R_xlen_t)d_->size(); }
    R_xlen_t length() const { return size(); }
    typename std::vector<T>::iterator begin() { return d_->begin(); }
    typename std::vector<T>::iterator end() { return d_->end(); }
    typename std::vector<T>::const_iterator begin() const { return d_->begin(); }
    typename std::vector<T>::const_iterator end() const { return d_->end(); }
    operator std::vector<T>() const { return *d_; }
    const std::vector<int>& dims() const { return *dim_; }
    void set_dims(const std::vector<int>& v) { *dim_ = v; }
    DimProxy<T> attr(const char*) { return DimProxy<T>(*this); }
    static bool is_na(double x) { return std::isnan(x); }
    static double get_na() { return NA_REAL; }
    T* raw() { return d_->data(); }
    Vector deep_copy() const {
        Vector r;
        *r.d_ = *d_;
        *r.dim_ = *dim_;
        return r;
    }
};

template <class T> DimProxy<T>& DimProxy<T>::operator=(const IntegerVector& d) {
    owner_.set_dims(std::vector<int>(d.begin(), d.end()));
    return *this;
}
template <class T> DimProxy<T>::operator IntegerVector() const { return IntegerVector(owner_.dims()); }

template <class T> class Matrix : public Vector<T> {
public:
    Matrix() : Vector<T>() {}
    Matrix(int r, int c) : Vector<T>((size_t)r * (size_t)c) { this->set_dims({r, c}); }
    Matrix(const Vector<T>& v) : Vector<T>(v) {}
    template <class U, class = typename std::enable_if<!std::is_same<U, T>::value>::type>
    Matrix(const Vector<U>& v) : Vector<T>(v) {}
    int nrow() const { return this->dims().size() > 0 ? this->dims()[0] : (int)this->size(); }
    int ncol() const { return this->dims().size() > 1 ? this->dims()[1] : 1; }
    T& operator()(int i, int j) { return (*this->d_)[(size_t)i + (size_t)nrow() * (size_t)j]; }
    const T& operator()(int i, int j) const { return (*this->d_)[(size_t)i + (size_t)nrow() * (size_t)j]; }
};
typedef Matrix<double> NumericMatrix;
typedef Matrix<int> IntegerMatrix;

class List;

// One list element: a numeric vector, an integer vector, or a nested list.
struct Element {
    enum Kind { NONE, NUM, INT, LIST } kind = NONE;
    NumericVector num;
    IntegerVector integer;
    std::shared_ptr<List> list;
    Element() {}
    Element(const NumericVector& v) : kind(NUM), num(v) {}
    Element(const IntegerVector& v) : kind(INT), integer(v) {}
    Element(const NumericMatrix& v) : kind(NUM), num(v) {}
    Element(const IntegerMatrix& v) : kind(INT), integer(v) {}
    Element(const std::vector<double>& v) : kind(NUM), num(v) {}
    Element(const std::vector<int>& v) : kind(INT), integer(v) {}
    Element(double v) : kind(NUM), num(NumericVector({v})) {}
    Element(int v) : kind(INT), integer(IntegerVector({v})) {}
    Element(bool v) : kind(INT), integer(IntegerVector({(int)v})) {}
    Element(const List& l);
    NumericVector as_num() const {
        if (kind == NUM) return num;
        if (kind == INT) return NumericVector(integer);
        throw std::runtime_error("Rcpp shim: list element is not numeric");
    }
    IntegerVector as_int() const {
        if (kind == INT) return integer;
        if (kind == NUM) return IntegerVector(num);
        throw std::runtime_error("Rcpp shim: list element is not integer");
    }
};

class ElementProxy {
    Element& e_;
public:
    explicit ElementProxy(Element& e) : e_(e) {}
    template <class V> ElementProxy& operator=(const V& v) { e_ = Element(v); return *this; }
    ElementProxy& operator=(const ElementProxy& o) { e_ = o.e_; return *this; }
    ElementProxy& operator=(const Element& o) { e_ = o; return *this; }
    const Element& element() const { return e_; }
    operator NumericVector() const { return e_.as_num(); }
    operator IntegerVector() const { return e_.as_int(); }
    operator NumericMatrix() const { return NumericMatrix(e_.as_num()); }
    operator IntegerMatrix() const { return IntegerMatrix(e_.as_int()); }
    operator std::vector<double>() const { return (std::vector<double>)e_.as_num(); }
    operator std::vector<int>() const { return (std::vector<int>)e_.as_int(); }
    operator double() const { return e_.as_num()[0]; }
    operator int() const { return e_.as_int()[0]; }
    operator bool() const { return e_.as_int()[0] != 0; }
    operator List() const;
};

struct NamedValue {
    std::string name;
    Element value;
};
struct NamedTag {
    std::string name;
    template <class V> NamedValue operator=(const V& v) const { return NamedValue{name, Element(v)}; }
    NamedValue operator=(const ElementProxy& p) const { return NamedValue{name, p.element()}; }
};
struct NamedMaker {
    NamedTag operator[](const char* n) const { return NamedTag{n}; }
};
static const NamedMaker _ = NamedMaker();

class List {
    std::shared_ptr<std::map<std::string, Element>> m_;
    std::shared_ptr<std::vector<std::string>> order_;
public:
    List() : m_(std::make_shared<std::map<std::string, Element>>()), order_(std::make_shared<std::vector<std::string>>()) {}
    ElementProxy operator[](const std::string& key) {
        if (m_->find(key) == m_->end()) order_->push_back(key);
        return ElementProxy((*m_)[key]);
    }
    ElementProxy operator[](const std::string& key) const {
        auto it = m_->find(key);
        if (it == m_->end()) throw std::runtime_error("Rcpp shim: no list element named '" + key + "'");
        return ElementProxy(it->second);
    }
    bool containsElementNamed(const char* key) const { return m_->find(key) != m_->end(); }
    const std::vector<std::string>& names() const { return *order_; }
    R_xlen_t size() const { return (R_xlen_t)m_->size(); }
    template <class... A> static List create(const A&... a) {
        List l;
        NamedValue vals[] = {a...};
        for (const NamedValue& nv : vals) l[nv.name] = nv.value;
        return l;
    }
    List deep_copy() const {
        List r;
        for (const std::string& k : *order_) {
            const Element& e = m_->at(k);
            if (e.kind == Element::NUM) r[k] = e.num.deep_copy();
            else if (e.kind == Element::INT) r[k] = e.integer.deep_copy();
            else r[k] = e;
        }
        return r;
    }
};
typedef List DataFrame;

inline Element::Element(const List& l) : kind(LIST), list(std::make_shared<List>(l)) {}
inline ElementProxy::operator List() const {
    if (e_.kind != Element::LIST) throw std::runtime_error("Rcpp shim: list element is not a list");
    return *e_.list;
}

inline NumericVector wrap(const std::vector<double>& v) { return NumericVector(v); }
inline IntegerVector wrap(const std::vector<int>& v) { return IntegerVector(v); }
inline NumericVector wrap(double v) { return NumericVector({v}); }
inline IntegerVector wrap(int v) { return IntegerVector({v}); }
inline NumericVector wrap(const NumericVector& v) { return v; }
inline IntegerVector wrap(const IntegerVector& v) { return v; }

template <class R> R as(const ElementProxy& p) { return (R)p; }
template <class R> R as(const NumericVector& v) { return (R)v; }
template <class R> R as(const IntegerVector& v) { return (R)v; }

template <class T> Vector<T> clone(const Vector<T>& v) { return v.deep_copy(); }
template <class T> Matrix<T> clone(const Matrix<T>& v) { return Matrix<T>(v.deep_copy()); }
inline List clone(const List& l) { return l.deep_copy(); }

[[noreturn]] inline void stop(const std::string& msg) { throw std::runtime_error(msg); }

} // namespace Rcpp

#endif
