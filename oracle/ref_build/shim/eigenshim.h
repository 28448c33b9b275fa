// ORACLE — TEST INFRASTRUCTURE ONLY.  Just enough of Eigen's dense surface for the reference's track_calibration.cc to
// compile unmodified: MatrixXd (dynamic, double) with setOnes / setIdentity / (i, j) / (i) / rows / cols / transpose /
// * + - / determinant, and JacobiSVD<MatrixXd> whose numerics are forwarded to the restated algorithm (orc_linalg.h
// svd3_jacobi).  Products accumulate in ascending inner index.  Not a copy of any Eigen header.
#pragma once
#include <cassert>
#include <cstddef>
#include <vector>

#include "../../orc_linalg.h"

namespace Eigen {

enum { ComputeThinU = 1, ComputeThinV = 2, ComputeFullU = 4, ComputeFullV = 8 };

class MatrixXd {
 public:
  MatrixXd() {}
  MatrixXd(int r, int c) : r_(r), c_(c), d_((size_t)r * c, 0.0) {}
  void setOnes(int r, int c) { r_ = r; c_ = c; d_.assign((size_t)r * c, 1.0); }
  void setIdentity(int r, int c) {
    r_ = r; c_ = c; d_.assign((size_t)r * c, 0.0);
    for (int i = 0; i < r && i < c; i++) (*this)(i, i) = 1.0;
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  double& operator()(int i, int j) { return d_[(size_t)i * c_ + j]; }
  double operator()(int i, int j) const { return d_[(size_t)i * c_ + j]; }
  double& operator()(int i) { return d_[(size_t)i]; }  // vectors only (N x 1 or 1 x N)
  double operator()(int i) const { return d_[(size_t)i]; }
  MatrixXd transpose() const {
    MatrixXd t(c_, r_);
    for (int i = 0; i < r_; i++)
      for (int j = 0; j < c_; j++) t(j, i) = (*this)(i, j);
    return t;
  }
  double determinant() const {
    assert(r_ == 3 && c_ == 3);
    const MatrixXd& m = *this;
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
  }

 private:
  int r_ = 0, c_ = 0;
  std::vector<double> d_;
};

inline MatrixXd operator*(const MatrixXd& a, const MatrixXd& b) {
  assert(a.cols() == b.rows());
  MatrixXd c(a.rows(), b.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < b.cols(); j++) {
      double s = a(i, 0) * b(0, j);
      for (int k = 1; k < a.cols(); k++) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}
inline MatrixXd operator-(const MatrixXd& a, const MatrixXd& b) {
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  MatrixXd c(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < a.cols(); j++) c(i, j) = a(i, j) - b(i, j);
  return c;
}
inline MatrixXd operator+(const MatrixXd& a, const MatrixXd& b) {
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  MatrixXd c(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < a.cols(); j++) c(i, j) = a(i, j) + b(i, j);
  return c;
}

template <class M>
class JacobiSVD {
 public:
  JacobiSVD(const M& h, unsigned) : u_(3, 3), v_(3, 3), s_(3, 1) {
    assert(h.rows() == 3 && h.cols() == 3);
    double H[9], U[9], S[3], V[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) H[i * 3 + j] = h(i, j);
    orc::svd3_jacobi(H, U, S, V);
    for (int i = 0; i < 3; i++) {
      s_(i, 0) = S[i];
      for (int j = 0; j < 3; j++) {
        u_(i, j) = U[i * 3 + j];
        v_(i, j) = V[i * 3 + j];
      }
    }
  }
  const M& matrixU() const { return u_; }
  const M& matrixV() const { return v_; }
  const M& singularValues() const { return s_; }

 private:
  M u_, v_, s_;
};

}  // namespace Eigen
