/* Stand-in for jrl-mal's <jrl/mal/matrixabstractlayer.hh> (jrl-mal >= 1.9.0 is an
 * un-vendored dependency of the reference, CMakeLists.txt:45).  Written for this repository;
 * it only provides the MAL_* macros used by the reference sources that oracle/Makefile
 * compiles where they lie (OptCholesky.cpp and PLDPSolver.cpp use none;
 * FootConstraintsAsLinearSystem.cpp and pgtypes.hh use MAL_MATRIX / MAL_VECTOR, *_RESIZE,
 * MAL_MATRIX_NB_ROWS, MAL_VECTOR_DIM and MAL_RET_A_by_B; PreviewControl.cpp and
 * OptimalControllerSolver.cpp add matrix products, sums, scalar scaling, transposes, FILL /
 * SET_IDENTITY, the raw data block handed to LAPACK and MAL_INVERSE).  Semantics are those of
 * the Boost uBLAS backend of jrl-mal: dense row-major double storage, resize keeps the old
 * contents; NEW elements are zero here (uBLAS leaves them uninitialised); a product element is
 * t = 0; t += a(i,k)*b(k,j) for k ascending (uBLAS matrix_matrix_prod / matrix_vector_prod1),
 * and `prod(A,x) + s*B` is evaluated element by element as (sum) + (s*B(i,j)), which is what
 * the uBLAS expression templates do.  MAL_INVERSE is LAPACK dgetrf_ + dgetri_ as in jrl-mal's
 * boost backend (the LAPACK symbols are resolved by oracle/ref_glue.cc at run time).
 * TEST INFRASTRUCTURE ONLY - nothing under oracle/ is linked into the product. */
#ifndef ORACLE_REF_SHIM_MAL_HH
#define ORACLE_REF_SHIM_MAL_HH
#include <vector>
#include <cstddef>

namespace oracle_mal {

template <typename T> class vector {
 public:
  vector() {}
  explicit vector(std::size_t n) : d_(n, T()) {}
  void resize(std::size_t n) { d_.resize(n, T()); }
  void resize(std::size_t n, bool) { d_.resize(n, T()); }     /* uBLAS resize(n, preserve) */
  void clear() { fill(T()); }                                /* uBLAS clear(): zero the elements */
  std::size_t size() const { return d_.size(); }
  T &operator()(std::size_t i) { return d_[i]; }
  const T &operator()(std::size_t i) const { return d_[i]; }
  T &operator[](std::size_t i) { return d_[i]; }
  const T &operator[](std::size_t i) const { return d_[i]; }
  T *data() { return d_.empty() ? 0 : &d_[0]; }
  void fill(T v) { for (std::size_t i = 0; i < d_.size(); ++i) d_[i] = v; }
 private:
  std::vector<T> d_;
};

template <typename T> class matrix {
  /* up to 16 elements live inside the object: the 3x3 / 3x1 / 1x1 temporaries of the preview recursion then cost no
   * heap allocation (uBLAS' unbounded_array would new[] each of them; the stand-in must not make the reference's
   * per-tick call slower than it is) */
  enum { INLINE = 16 };
 public:
  matrix() : r_(0), c_(0), p_(in_) {}
  matrix(std::size_t r, std::size_t c) : r_(0), c_(0), p_(in_) { alloc(r, c); }
  matrix(const matrix &o) : r_(0), c_(0), p_(in_) { *this = o; }
  matrix &operator=(const matrix &o)
  {
    if (this != &o) {
      if (o.r_ * o.c_ != r_ * c_) alloc_raw(o.r_, o.c_);
      r_ = o.r_; c_ = o.c_;
      for (std::size_t i = 0; i < r_ * c_; ++i) p_[i] = o.p_[i];
    }
    return *this;
  }
  void resize(std::size_t r, std::size_t c)
  {
    if (r == r_ && c == c_) return;
    matrix n(r, c);
    for (std::size_t i = 0; i < r && i < r_; ++i)
      for (std::size_t j = 0; j < c && j < c_; ++j) n(i, j) = (*this)(i, j);
    *this = n;
  }
  std::size_t size1() const { return r_; }
  std::size_t size2() const { return c_; }
  T &operator()(std::size_t i, std::size_t j) { return p_[i * c_ + j]; }
  const T &operator()(std::size_t i, std::size_t j) const { return p_[i * c_ + j]; }
  T *data() { return r_ * c_ ? p_ : 0; }
  void fill(T v) { for (std::size_t i = 0; i < r_ * c_; ++i) p_[i] = v; }
  void set_identity()
  {
    for (std::size_t i = 0; i < r_; ++i)
      for (std::size_t j = 0; j < c_; ++j) (*this)(i, j) = (i == j) ? T(1) : T(0);
  }
 private:
  void alloc_raw(std::size_t r, std::size_t c)
  {
    if (r * c <= INLINE) { heap_.clear(); p_ = in_; }
    else { heap_.assign(r * c, T()); p_ = &heap_[0]; }
  }
  void alloc(std::size_t r, std::size_t c)
  {
    alloc_raw(r, c);
    r_ = r; c_ = c;
    for (std::size_t i = 0; i < r * c; ++i) p_[i] = T();
  }
  std::size_t r_, c_;
  T in_[INLINE];
  std::vector<T> heap_;
  T *p_;
};

/* prod(matrix, matrix): element (i,j) = 0 + sum_k a(i,k)*b(k,j), k ascending */
template <typename T> matrix<T> prod(const matrix<T> &A, const matrix<T> &B)
{
  matrix<T> Cm(A.size1(), B.size2());
  for (std::size_t i = 0; i < A.size1(); ++i)
    for (std::size_t j = 0; j < B.size2(); ++j) {
      T t = T();
      for (std::size_t k = 0; k < A.size2(); ++k) t += A(i, k) * B(k, j);
      Cm(i, j) = t;
    }
  return Cm;
}
template <typename T> matrix<T> trans(const matrix<T> &A)
{
  matrix<T> R(A.size2(), A.size1());
  for (std::size_t i = 0; i < A.size1(); ++i)
    for (std::size_t j = 0; j < A.size2(); ++j) R(j, i) = A(i, j);
  return R;
}
template <typename T> matrix<T> operator+(const matrix<T> &A, const matrix<T> &B)
{
  matrix<T> R(A.size1(), A.size2());
  for (std::size_t i = 0; i < A.size1(); ++i)
    for (std::size_t j = 0; j < A.size2(); ++j) R(i, j) = A(i, j) + B(i, j);
  return R;
}
template <typename T> matrix<T> operator-(const matrix<T> &A, const matrix<T> &B)
{
  matrix<T> R(A.size1(), A.size2());
  for (std::size_t i = 0; i < A.size1(); ++i)
    for (std::size_t j = 0; j < A.size2(); ++j) R(i, j) = A(i, j) - B(i, j);
  return R;
}
template <typename T> matrix<T> operator-(const matrix<T> &A)
{
  matrix<T> R(A.size1(), A.size2());
  for (std::size_t i = 0; i < A.size1(); ++i)
    for (std::size_t j = 0; j < A.size2(); ++j) R(i, j) = -A(i, j);
  return R;
}
template <typename T> matrix<T> operator*(const matrix<T> &A, T s)
{
  matrix<T> R(A.size1(), A.size2());
  for (std::size_t i = 0; i < A.size1(); ++i)
    for (std::size_t j = 0; j < A.size2(); ++j) R(i, j) = A(i, j) * s;
  return R;
}
template <typename T> matrix<T> operator*(T s, const matrix<T> &A)
{
  matrix<T> R(A.size1(), A.size2());
  for (std::size_t i = 0; i < A.size1(); ++i)
    for (std::size_t j = 0; j < A.size2(); ++j) R(i, j) = s * A(i, j);
  return R;
}
template <typename S, typename T> S &operator<<(S &os, const matrix<T> &A)
{
  os << "[" << A.size1() << "," << A.size2() << "](";
  for (std::size_t i = 0; i < A.size1(); ++i) {
    os << (i ? ",(" : "(");
    for (std::size_t j = 0; j < A.size2(); ++j) os << (j ? "," : "") << A(i, j);
    os << ")";
  }
  os << ")";
  return os;
}
/* inverse of a square matrix by LU (dgetrf_ / dgetri_); defined in oracle/ref_glue.cc */
void lapack_inverse(const matrix<double> &A, matrix<double> &invA);

/* prod(matrix, vector): plain row-by-row accumulation, as uBLAS' dense matrix-vector product */
template <typename T> vector<T> prod(const matrix<T> &A, const vector<T> &x)
{
  vector<T> y(A.size1());
  for (std::size_t i = 0; i < A.size1(); ++i) {
    T t = T();
    for (std::size_t j = 0; j < A.size2(); ++j) t += A(i, j) * x(j);
    y(i) = t;
  }
  return y;
}

/* compound assignments and vector sums used by ZMPQPWithConstraint.cpp (:968, :1064, :1252): element by element */
template <typename T> matrix<T> &operator+=(matrix<T> &A, const matrix<T> &B)
{
  for (std::size_t i = 0; i < A.size1(); ++i)
    for (std::size_t j = 0; j < A.size2(); ++j) A(i, j) += B(i, j);
  return A;
}
template <typename T> vector<T> &operator-=(vector<T> &a, const vector<T> &b)
{
  for (std::size_t i = 0; i < a.size(); ++i) a(i) -= b(i);
  return a;
}
template <typename T> vector<T> operator+(const vector<T> &a, const vector<T> &b)
{
  vector<T> r(a.size());
  for (std::size_t i = 0; i < a.size(); ++i) r(i) = a(i) + b(i);
  return r;
}

/* jrl-mal's "small" fixed-size types (matrixabstractlayersmall*.hh): a 3-vector and a 4x4 matrix, only element access
 * and assignment are used by the sources compiled here (ZMPPreviewControlWithMultiBodyZMP.cpp) */
template <typename T> struct vec3 {
  T v[3];
  vec3() { v[0] = v[1] = v[2] = T(); }
  T &operator()(std::size_t i) { return v[i]; }
  const T &operator()(std::size_t i) const { return v[i]; }
  T &operator[](std::size_t i) { return v[i]; }
  const T &operator[](std::size_t i) const { return v[i]; }
};
template <typename T> struct mat4 {
  T m[16];
  mat4() { for (int i = 0; i < 16; ++i) m[i] = T(); }
  T &operator()(std::size_t i, std::size_t j) { return m[4 * i + j]; }
  const T &operator()(std::size_t i, std::size_t j) const { return m[4 * i + j]; }
};

}  // namespace oracle_mal

#define MAL_VECTOR_TYPE(type) oracle_mal::vector<type>
#define MAL_S3_VECTOR(name, type) oracle_mal::vec3<type> name
#define MAL_S3_VECTOR_TYPE(type) oracle_mal::vec3<type>
#define MAL_S4x4_MATRIX_TYPE(type) oracle_mal::mat4<type>
#define MAL_S4x4_MATRIX(name, type) oracle_mal::mat4<type> name
#define MAL_S4x4_MATRIX_ACCESS_I_J(name, i, j) name(i, j)
#define MAL_VECTOR(name, type) oracle_mal::vector<type> name
#define MAL_VECTOR_DIM(name, type, n) oracle_mal::vector<type> name(n)
#define MAL_VECTOR_RESIZE(name, n) name.resize(n)
#define MAL_VECTOR_SIZE(name) name.size()
#define MAL_MATRIX(name, type) oracle_mal::matrix<type> name
#define MAL_MATRIX_DIM(name, type, r, c) oracle_mal::matrix<type> name(r, c)
#define MAL_MATRIX_RESIZE(name, r, c) name.resize(r, c)
#define MAL_MATRIX_NB_ROWS(name) name.size1()
#define MAL_MATRIX_NB_COLS(name) name.size2()
#define MAL_RET_A_by_B(A, B) oracle_mal::prod(A, B)
#define MAL_C_eq_A_by_B(C, A, B) (C) = oracle_mal::prod(A, B)   /* uBLAS: C = prod(A, B) through a temporary (aliasing safe) */
#define MAL_MATRIX_TYPE(type) oracle_mal::matrix<type>
#define MAL_MATRIX_FILL(name, v) name.fill(v)
#define MAL_VECTOR_FILL(name, v) name.fill(v)
#define MAL_MATRIX_SET_IDENTITY(name) name.set_identity()
#define MAL_RET_TRANSPOSE(name) oracle_mal::trans(name)
#define MAL_RET_MATRIX_DATABLOCK(name) name.data()
#define MAL_RET_VECTOR_DATABLOCK(name) name.data()
#define MAL_INVERSE(name, inv, type) oracle_mal::lapack_inverse(name, inv)
#endif
