/* TEST INFRASTRUCTURE ONLY (oracle/Makefile).  Stand-in for the two Boost uBLAS headers privatepgtypes.hh includes: the sparse
 * matrix / vector members it declares belong to the Herdt-era QP structs (linear_inequality_t, solution_t), which none of the
 * reference sources compiled into oracle/_ref touch; the types only have to exist, be resizable and be clearable. */
#ifndef ORACLE_REF_SHIM_UBLAS_MATRIX_SPARSE_HPP
#define ORACLE_REF_SHIM_UBLAS_MATRIX_SPARSE_HPP
#include <cstddef>
#include <vector>
namespace boost { namespace numeric { namespace ublas {
struct row_major {};
template <class T, class L = row_major> class compressed_matrix {
 public:
  compressed_matrix() : r_(0), c_(0) {}
  void resize(std::size_t r, std::size_t c, bool = true) { r_ = r; c_ = c; }
  void clear() {}
  std::size_t size1() const { return r_; }
  std::size_t size2() const { return c_; }
 private:
  std::size_t r_, c_;
};
template <class T> class vector {
 public:
  vector() {}
  explicit vector(std::size_t n) : d_(n) {}
  void resize(std::size_t n, bool = true) { d_.resize(n); }
  void clear() { for (std::size_t i = 0; i < d_.size(); ++i) d_[i] = T(); }
  std::size_t size() const { return d_.size(); }
  T &operator()(std::size_t i) { return d_[i]; }
  const T &operator()(std::size_t i) const { return d_[i]; }
  T &operator[](std::size_t i) { return d_[i]; }
  const T &operator[](std::size_t i) const { return d_[i]; }
 private:
  std::vector<T> d_;
};
}}}
namespace boost_ublas = boost::numeric::ublas;
#endif
