/* oracle/oracle_wieber.cpp - TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the Wieber2006 generator: ZMPQPWithConstraint::BuildMatricesPxPu and
 * BuildZMPTrajectoryFromFootTrajectory (src/ZMPRefTrajectoryGeneration/ZMPQPWithConstraint.cpp:504-663, :665-1338).
 * The support polygons (BuildLinearConstraintInequalities / ComputeLinearSystem, :94-502) are the same code as
 * FootConstraintsAsLinearSystem's (A and B identical; no centre, no similar flags; StartingTime / EndingTime are the
 * `time` field of the left-foot samples instead of an accumulated clock) and are taken from oracle_fcals_build
 * (oracle_dimitrov.cpp).  The QP is solved by the reference's own ql0001_ (function pointer handed over by the test from
 * oracle/_ref) or, failing that, by oracle_qp_solve.  Matrix products follow the order of the uBLAS expressions the
 * reference uses (sum over k ascending from 0), so that with the reference's ql0001_ the outputs are bitwise those of the
 * reference object code (tests/test_wieber.py).
 */
#include <cmath>
#include <cstring>
#include <vector>
#include "../include/walkgen_b200.h"

extern "C" int oracle_fcals_build(int n, const double *left, const double *right, const int *step_type, const double *time,
                                  double sole_length, double sole_width, double cx, double cy, int cap, wg_lci *out, int merge);
extern "C" int oracle_qp_solve(int n, int m, const double *C, const double *dvec, const double *A, const double *b,
                               double *x_out, double *u_out, int *iterations);
extern "C" int oracle_qp_solve_ld(int n, int m, const double *C, const double *dvec, const double *A, const double *b,
                                  double *x_out, double *u_out, int *iterations);

namespace {
typedef int (*ql0001_fn)(int *, int *, int *, int *, int *, int *, double *, double *, double *, double *, double *,
                         double *, double *, double *, int *, int *, int *, double *, int *, int *, int *, double *);
ql0001_fn g_ql = nullptr;
double g_boost = 0.0; /* multiple of I added to the Hessian for the textbook solvers: the diag QLD adds (qld.cpp:809-918) */
int g_solver = 0;   /* 0: the reference's ql0001_ when set, else textbook double; 1: textbook double; 2: textbook long double */

struct Mat {
  int r, c;
  std::vector<double> a;
  Mat(int r_ = 0, int c_ = 0) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double &operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};
Mat prod(const Mat &A, const Mat &B)
{
  Mat C(A.r, B.c);
  for (int i = 0; i < A.r; ++i)
    for (int j = 0; j < B.c; ++j) {
      double t = 0.0;
      for (int k = 0; k < A.c; ++k) t += A(i, k) * B(k, j);
      C(i, j) = t;
    }
  return C;
}
Mat trans(const Mat &A)
{
  Mat C(A.c, A.r);
  for (int i = 0; i < A.r; ++i)
    for (int j = 0; j < A.c; ++j) C(j, i) = A(i, j);
  return C;
}
}  // namespace

extern "C" {

void oracle_wieber_set_qld(void *fn) { g_ql = (ql0001_fn)fn; }
void oracle_wieber_set_solver(int which) { g_solver = which; }
void oracle_wieber_set_boost(double diag) { g_boost = diag; }

/* The constant matrices of :700-770, :905-990 for tests of the product's constants: C (2N x 2N, symmetric),
 * OptB (2N x 6), OptC (2N x 2N), all row-major. */
void oracle_wieber_constants(int N, double T, double alpha, double beta, double *Cout, double *OptBout, double *OptCout)
{
  const int n = 2 * N;
  Mat PPu(n, n), VPu(n, n), PPx(n, 6), VPx(n, 6);
  for (int i = 0; i < N; ++i) {
    VPx(i, 1) = 1.0; VPx(i, 2) = (i + 1) * T;
    VPx(i + N, 4) = 1.0; VPx(i + N, 5) = (i + 1) * T;
    PPx(i, 0) = 1.0; PPx(i, 1) = (i + 1) * T; PPx(i, 2) = (i + 1) * (i + 1) * T * T * 0.5;
    PPx(i + N, 3) = 1.0; PPx(i + N, 4) = (i + 1) * T; PPx(i + N, 5) = (i + 1) * (i + 1) * T * T * 0.5;
    for (int j = 0; j <= i; ++j) {
      VPu(i, j) = VPu(i + N, j + N) = (2 * (i - j) + 1) * T * T * 0.5;
      PPu(i, j) = PPu(i + N, j + N) = (1 + 3 * (i - j) + 3 * (i - j) * (i - j)) * T * T * T / 6.0;
    }
  }
  Mat l1 = prod(trans(PPu), PPu), l2 = prod(trans(VPu), VPu);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Cout[(size_t)i * n + j] = beta * l1(i, j) + alpha * l2(i, j);
  Mat b1 = prod(trans(PPu), PPx), b2 = prod(trans(VPu), VPx);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 6; ++j) {
      double v = alpha * b2(i, j);
      v += beta * b1(i, j);
      OptBout[(size_t)i * 6 + j] = v;
    }
  Mat tp = trans(PPu);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) OptCout[(size_t)i * n + j] = beta * tp(i, j);
}

/* BuildMatricesPxPu (:504-663).  lci: polygons (A, B, t_start, t_end, rows).  Px [>= 8N+1], Pu [(8N+1) * 2N] column-major
 * with leading dimension NbOfConstraints + 1 (as the reference lays it out).  Returns 0, or -1 ("HERE 3"). */
int oracle_wieber_build(int N, double T, double StartingTime, int np, const wg_lci *lci, double ComHeight, const double *xk,
                        double *Px, double *Pu, int *nb_out)
{
  std::memset(Pu, 0, sizeof(double) * (size_t)(8 * N + 1) * 2 * N);
  int it = 0;
  while (it < np) {
    if (StartingTime >= lci[it].t_start && StartingTime <= lci[it].t_end) break;
    ++it;
  }
  const int store = it;
  if (it == np) return -1;
  int Index = 0;
  for (int i = 0; i < N; ++i) {
    const double ltime = StartingTime + i * T;
    if (ltime > lci[it].t_end) ++it;
    if (it == np) break;
    Index += lci[it].rows;
  }
  const int Nb = Index;
  *nb_out = Nb;
  it = store;
  Index = 0;
  for (int i = 0; i < N; ++i) {
    const double ltime = StartingTime + i * T;
    if (ltime > lci[it].t_end) ++it;
    if (it >= np) return -2;                 /* the reference dereferences end() here */
    for (int j = 0; j < lci[it].rows; ++j) {
      Px[Index] = (xk[0] + xk[1] * T * (i + 1) + xk[2] * ((i + 1) * (i + 1) * T * T / 2 - ComHeight / 9.81)) * lci[it].A[j][0] +
                  (xk[3] + xk[4] * T * (i + 1) + xk[5] * ((i + 1) * (i + 1) * T * T / 2 - ComHeight / 9.81)) * lci[it].A[j][1] +
                  lci[it].B[j];
      for (int k = 0; k <= i; ++k) {
        Pu[Index + (size_t)k * (Nb + 1)] = lci[it].A[j][0] * ((1 + 3 * (i - k) + 3 * (i - k) * (i - k)) * T * T * T / 6.0 - T * ComHeight / 9.81);
        Pu[Index + (size_t)(k + N) * (Nb + 1)] = lci[it].A[j][1] * ((1 + 3 * (i - k) + 3 * (i - k) * (i - k)) * T * T * T / 6.0 - T * ComHeight / 9.81);
      }
      ++Index;
    }
  }
  return 0;
}

/* BuildZMPTrajectoryFromFootTrajectory (:665-1338).  feet [n][4] = x, y, z, theta(deg); step_type / time: the left
 * foot's; zmp [n][3] = px, py, theta in/out; com [n][7] = x[0..2], y[0..2], yaw out (rows the loop does not reach stay
 * untouched).  Returns the number of QP periods solved, or -(1 + period) when the reference would `return -1` there
 * (ifail != 0, violated constraint, no polygon).  info [periods][3] (may be NULL) = m, ifail, active rows. */
long oracle_wieber_run(long n, const double *left, const double *right, const int *step_type, const double *time,
                       double *zmp, double *com, double sole_length, double sole_width, double cx, double cy, double T, int N,
                       double Ts, long max_periods, int *info)
{
  const double ComHeight = 0.80, alpha = 200.0, beta = 1000.0;
  std::vector<wg_lci> lci(4096);
  const int np = oracle_fcals_build((int)n, left, right, step_type, time, sole_length, sole_width, cx, cy, (int)lci.size(),
                                    lci.data(), 0);
  if (np <= 0 || np > (int)lci.size()) return -1000000;
  const int nv = 2 * N;
  std::vector<double> Cm((size_t)nv * nv), OptB((size_t)nv * 6), OptC((size_t)nv * nv);
  oracle_wieber_constants(N, T, alpha, beta, Cm.data(), OptB.data(), OptC.data());
  std::vector<double> Ccm((size_t)nv * nv);
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) Ccm[(size_t)j * nv + i] = Cm[(size_t)i * nv + j];
  double xk[6] = {0, 0, 0, 0, 0, 0};
  std::vector<double> Px(8 * N + 1), Pu((size_t)(8 * N + 1) * nv), D(nv), ZMPRef(nv), X(nv), XL(nv, -1e8), XU(nv, 1e8);
  const int interval = (int)(T / Ts);
  const double mA[3][3] = {{1.0, T, T * T / 2.0}, {0.0, 1.0, T}, {0.0, 0.0, 1.0}};
  const double mB[3] = {T * T * T / 6.0, T * T / 2.0, T};
  const double mC[3] = {1.0, 0.0, -ComHeight / 9.81};
  long li = 0;
  for (double StartingTime = 0.0; StartingTime < lci[np - 1].t_end - N * T; StartingTime += T, ++li) {
    if (max_periods > 0 && li >= max_periods) break;
    int m = 0;
    if (oracle_wieber_build(N, T, StartingTime, np, lci.data(), ComHeight, xk, Px.data(), Pu.data(), &m) != 0) return -(1 + li);
    for (int i = 0; i < N; ++i) {
      ZMPRef[i] = zmp[3 * (li * interval + (long)i * interval)];
      ZMPRef[i + N] = zmp[3 * (li * interval + (long)i * interval) + 1];
    }
    for (int i = 0; i < nv; ++i) {
      double t1 = 0.0, t2 = 0.0;
      for (int k = 0; k < nv; ++k) t1 += OptC[(size_t)i * nv + k] * ZMPRef[k];
      for (int k = 0; k < 6; ++k) t2 += OptB[(size_t)i * 6 + k] * xk[k];
      D[i] = t2 - t1;
    }
    std::fill(X.begin(), X.end(), 0.0);
    int ifail = 0, nact = 0;
    if (g_ql && g_solver == 0) {
      int mm = m, me = 0, mmax = m + 1, nn = nv, nmax = nv, mnn = m + 2 * nv, iout = 0, iprint = 1;
      int lwar = 3 * nmax * nmax / 2 + 10 * nmax + 2 * mmax + 20000, liwar = nv;
      std::vector<double> war(lwar), U(mnn), Cc(Ccm), Dd(D);
      std::vector<int> iwar(liwar + 8);
      iwar[0] = 1;
      double eps = 1e-8;
      g_ql(&mm, &me, &mmax, &nn, &nmax, &mnn, Cc.data(), Dd.data(), Pu.data(), Px.data(), XL.data(), XU.data(), X.data(), U.data(),
           &iout, &ifail, &iprint, war.data(), &lwar, iwar.data(), &liwar, &eps);
      for (int i = 0; i < m; ++i) nact += (U[i] != 0.0);
    } else {
      std::vector<double> Arm((size_t)m * nv), U(m), Cb(Cm);
      for (int i = 0; i < nv; ++i) Cb[(size_t)i * nv + i] += g_boost;
      for (int r = 0; r < m; ++r)
        for (int j = 0; j < nv; ++j) Arm[(size_t)r * nv + j] = Pu[r + (size_t)j * (m + 1)];
      ifail = (g_solver == 2 ? oracle_qp_solve_ld : oracle_qp_solve)(nv, m, Cb.data(), D.data(), Arm.data(), Px.data(), X.data(), U.data(), nullptr);
      for (int i = 0; i < m; ++i) nact += (U[i] != 0.0);
    }
    if (info) { info[3 * li] = m; info[3 * li + 1] = ifail; info[3 * li + 2] = nact; }
    if (ifail != 0) return -(1 + li);
    for (int i = 0; i < m; ++i) {           /* vnlValConstraint = vnlPu vnlX + vnlPx, :1070-1105 */
      double t = 0.0;
      for (int j = 0; j < nv; ++j) t += Pu[i + (size_t)j * (m + 1)] * X[j];
      if (t + Px[i] < -1e-8) return -(1 + li);
    }
    const double Buk[6] = {X[0] * mB[0], X[0] * mB[1], X[0] * mB[2], X[N] * mB[0], X[N] * mB[1], X[N] * mB[2]};
    for (int lk = 0; lk < interval; ++lk) {
      const long row = li * interval + lk;
      if (row >= n) break;
      const double s = (lk + 1) * Ts;
      double *c = com + 7 * row;
      c[0] = xk[0] + s * xk[1] + 0.5 * s * s * xk[2] + s * s * s * X[0] / 6.0;
      c[1] = xk[1] + s * xk[2] + 0.5 * s * s * X[0];
      c[2] = xk[2] + s * X[0];
      c[3] = xk[3] + s * xk[4] + 0.5 * s * s * xk[5] + s * s * s * X[N] / 6.0;
      c[4] = xk[4] + s * xk[5] + 0.5 * s * s * X[N];
      c[5] = xk[5] + s * X[N];
      c[6] = zmp[3 * row + 2];
      zmp[3 * row] = mC[0] * c[0] + mC[1] * c[1] + mC[2] * c[2];
      zmp[3 * row + 1] = mC[0] * c[3] + mC[1] * c[4] + mC[2] * c[5];
    }
    double nx[6];
    for (int a = 0; a < 2; ++a)
      for (int i = 0; i < 3; ++i) {
        double t = 0.0;
        for (int j = 0; j < 6; ++j) {                 /* prod(m_A 6 x 6, xk): the zero blocks are summed too */
          const double aij = (j / 3 == a) ? mA[i][j % 3] : 0.0;
          t += aij * xk[j];
        }
        nx[3 * a + i] = t + Buk[3 * a + i];
      }
    std::memcpy(xk, nx, sizeof nx);
  }
  return li;
}

}  /* extern "C" */
