/* oracle/oracle_qp.cpp - TEST INFRASTRUCTURE ONLY.
 *
 * Dense strictly-convex QP solver used by the oracle when it needs a QP solution:
 *      min 1/2 x'Cx + d'x   s.t.  A x + b >= 0
 * which is the problem the reference hands to ql0001_ (src/Mathematics/qld.cpp:378-612; Powell's
 * ZQPCVX dual active-set method).  The reference's own QLD object code is built into
 * oracle/_ref/libwalkgen_ref.so and is the primary pin; this file is an independent textbook
 * statement of the Goldfarb-Idnani dual active-set method (Math. Prog. 27, 1983) - the same family
 * as ZQPCVX - used (a) to cross-check QLD in tests and (b) as the oracle's solver on a machine
 * where oracle/_ref was never built.  Because the QP is strictly convex its minimiser and (in the
 * non-degenerate case) the set of positive multipliers are unique, so any exact solver is a valid
 * oracle for "identical optimal active sets".
 *
 * Dense, unoptimised, row-major.  A is m x n row-major here (the QLD calling convention with its
 * column-major array and dummy row is handled by the callers).
 */
#include <cmath>
#include <cstring>
#include <vector>
#include <limits>
#include <algorithm>

namespace {

template <typename Rr> struct GI {
  typedef Rr real;
  int n, m;
  std::vector<real> J, R;  /* J n x n (columns rotate), R n x n upper triangular (q x q used) */
  std::vector<real> x, u, s, d, z, r;
  std::vector<int> act;      /* active constraint indices, size q */
  std::vector<char> is_act;
  int q = 0;

  real &Jm(int i, int j) { return J[(size_t)i * n + j]; }
  real &Rm(int i, int j) { return R[(size_t)i * n + j]; }

  static void givens(real a, real b, real &c, real &s)
  {
    if (b == 0.0) { c = 1.0; s = 0.0; return; }
    real h = std::sqrt(a * a + b * b);
    c = a / h; s = b / h;
  }

  /* d = J' a ; z = J2 d2 ; r = R^-1 d1 */
  void compute_d(const real *a)
  {
    for (int j = 0; j < n; ++j) {
      real t = 0;
      for (int i = 0; i < n; ++i) t += Jm(i, j) * a[i];
      d[j] = t;
    }
  }
  void update_z()
  {
    for (int i = 0; i < n; ++i) {
      real t = 0;
      for (int j = q; j < n; ++j) t += Jm(i, j) * d[j];
      z[i] = t;
    }
  }
  void update_r()
  {
    for (int i = q - 1; i >= 0; --i) {
      real t = d[i];
      for (int j = i + 1; j < q; ++j) t -= Rm(i, j) * r[j];
      r[i] = t / Rm(i, i);
    }
  }
  bool add_constraint()
  {
    /* rotate d[q+1..n-1] into d[q] */
    for (int j = n - 1; j > q; --j) {
      real c, s;
      givens(d[j - 1], d[j], c, s);
      if (s == 0.0 && c == 1.0) continue;
      real dj1 = c * d[j - 1] + s * d[j];
      d[j] = 0.0;
      d[j - 1] = dj1;
      for (int i = 0; i < n; ++i) {
        real a = Jm(i, j - 1), b = Jm(i, j);
        Jm(i, j - 1) = c * a + s * b;
        Jm(i, j) = -s * a + c * b;
      }
    }
    for (int i = 0; i <= q; ++i) Rm(i, q) = d[i];
    if (std::fabs(d[q]) <= (sizeof(real) > 8 ? 1e-17 : 1e-14) * std::fabs(Rm(0, 0) == 0 ? real(1.0) : Rm(0, 0))) return false;
    ++q;
    return true;
  }
  void drop_constraint(int l)
  {
    /* remove column l of R, restore triangular form */
    for (int j = l; j < q - 1; ++j) {
      for (int i = 0; i <= j + 1; ++i) Rm(i, j) = Rm(i, j + 1);
      act[j] = act[j + 1];
      u[j] = u[j + 1];
    }
    --q;
    for (int i = 0; i <= q; ++i) Rm(i, q) = 0.0;
    for (int j = l; j < q; ++j) {
      real c, s;
      givens(Rm(j, j), Rm(j + 1, j), c, s);
      for (int k = j; k < q; ++k) {
        real a = Rm(j, k), b = Rm(j + 1, k);
        Rm(j, k) = c * a + s * b;
        Rm(j + 1, k) = -s * a + c * b;
      }
      for (int i = 0; i < n; ++i) {
        real a = Jm(i, j), b = Jm(i, j + 1);
        Jm(i, j) = c * a + s * b;
        Jm(i, j + 1) = -s * a + c * b;
      }
    }
  }
};

} // namespace

/* Returns 0 on success, 1 if the iteration limit was hit, 2 if C is not positive definite,
 * 3 if the constraints are inconsistent.  x[n]; u[m] multipliers (0 for inactive rows);
 * iterations (may be NULL) counts constraint additions + removals.  real: the arithmetic (double, or long double for the
 * extended-precision oracle of badly conditioned problems). */
template <typename real>
int qp_solve_t(int n, int m, const double *C, const double *dvec, const double *A, const double *b,
               double *x_out, double *u_out, int *iterations)
{
  GI<real> g;
  g.n = n; g.m = m;
  g.J.assign((size_t)n * n, 0.0); g.R.assign((size_t)n * n, 0.0);
  g.x.assign(n, 0.0); g.u.assign(n + 1, 0.0); g.s.assign(m, 0.0);
  g.d.assign(n, 0.0); g.z.assign(n, 0.0); g.r.assign(n + 1, 0.0);
  g.act.assign(n + 1, -1); g.is_act.assign(m, 0);

  /* Cholesky C = L L' */
  std::vector<real> L((size_t)n * n, 0.0);
  for (int j = 0; j < n; ++j) {
    real t = C[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) t -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
    if (!(t > 0.0)) return 2;
    L[(size_t)j * n + j] = std::sqrt(t);
    for (int i = j + 1; i < n; ++i) {
      real v = C[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) v -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
      L[(size_t)i * n + j] = v / L[(size_t)j * n + j];
    }
  }
  /* J = L^-T : column j of J solves L' J(:,j) = e_j */
  for (int j = 0; j < n; ++j) {
    for (int i = n - 1; i >= 0; --i) {
      real t = (i == j) ? 1.0 : 0.0;
      for (int k = i + 1; k < n; ++k) t -= L[(size_t)k * n + i] * g.Jm(k, j);
      g.Jm(i, j) = t / L[(size_t)i * n + i];
    }
  }
  /* x = -C^-1 d = -J J' d */
  {
    std::vector<real> t(n);
    for (int j = 0; j < n; ++j) {
      real v = 0;
      for (int i = 0; i < n; ++i) v += g.Jm(i, j) * dvec[i];
      t[j] = v;
    }
    for (int i = 0; i < n; ++i) {
      real v = 0;
      for (int j = 0; j < n; ++j) v += g.Jm(i, j) * t[j];
      g.x[i] = -v;
    }
  }
  real scale = 0.0;
  std::vector<real> rown(m, 1.0);
  for (int i = 0; i < m; ++i) {
    real nr = 0;
    for (int j = 0; j < n; ++j) nr += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    scale = std::max(scale, std::sqrt(nr));
    rown[i] = nr > 0 ? std::sqrt(nr) : real(1.0);
  }
  const real tol = (sizeof(real) > 8 ? 1e-15 : 1e-12) * (scale > 0 ? scale : real(1.0));
  int iters = 0;
  const int maxit = 40 * (m + n);
  const real inf = std::numeric_limits<real>::infinity();
  std::vector<real> apv(n);

  for (;;) {
    /* most violated constraint */
    int p = -1;
    real smin = -tol;
    for (int i = 0; i < m; ++i) {
      if (g.is_act[i]) continue;
      real v = b[i];
      for (int j = 0; j < n; ++j) v += (real)A[(size_t)i * n + j] * g.x[j];
      g.s[i] = v;
      if (v < smin) { smin = v; p = i; }
    }
    if (p < 0) break;
    for (int j = 0; j < n; ++j) apv[j] = A[(size_t)p * n + j];
    const real *ap = apv.data();
    real up = 0.0;
    real sp = g.s[p];
    for (;;) {
      if (++iters > maxit) return 1;
      g.compute_d(ap);
      g.update_z();
      g.update_r();
      real t1 = inf; int l = -1;
      for (int k = 0; k < g.q; ++k)
        if (g.r[k] > 0.0) {
          real v = g.u[k] / g.r[k];
          if (v < t1) { t1 = v; l = k; }
        }
      real zz = 0, za = 0;
      for (int i = 0; i < n; ++i) { zz += g.z[i] * g.z[i]; za += g.z[i] * ap[i]; }
      real t2 = (zz <= (sizeof(real) > 8 ? 1e-34 : 1e-28) * scale * scale || za <= 0.0) ? inf : -sp / za;
      real t = std::min(t1, t2);
      if (t == inf) return 3;
      if (t2 == inf) { /* dual step only */
        for (int k = 0; k < g.q; ++k) g.u[k] -= t * g.r[k];
        up += t;
        g.is_act[g.act[l]] = 0;
        g.drop_constraint(l);
        continue;
      }
      for (int i = 0; i < n; ++i) g.x[i] += t * g.z[i];
      for (int k = 0; k < g.q; ++k) g.u[k] -= t * g.r[k];
      up += t;
      if (t == t2) { /* full step: constraint p becomes active */
        g.compute_d(ap);
        g.act[g.q] = p;
        g.u[g.q] = up;
        g.is_act[p] = 1;
        if (!g.add_constraint()) return 3;
        break;
      }
      /* partial step: drop l, recompute the violation of p */
      g.is_act[g.act[l]] = 0;
      g.drop_constraint(l);
      sp = b[p];
      for (int j = 0; j < n; ++j) sp += ap[j] * g.x[j];
    }
  }
  for (int i = 0; i < n; ++i) x_out[i] = (double)g.x[i];
  if (u_out) {
    for (int i = 0; i < m; ++i) u_out[i] = 0.0;
    for (int k = 0; k < g.q; ++k) u_out[g.act[k]] = (double)g.u[k];
  }
  if (iterations) *iterations = iters;
  return 0;
}

extern "C" {

int oracle_qp_solve(int n, int m, const double *C, const double *dvec, const double *A, const double *b,
                    double *x_out, double *u_out, int *iterations)
{
  return qp_solve_t<double>(n, m, C, dvec, A, b, x_out, u_out, iterations);
}
/* the same method in x87 extended precision (64-bit mantissa): the reference point for problems whose conditioning
 * (cond(C) ~ 5e11 for Wieber2006) puts a double-precision solution - the reference's ql0001_ included - 1e-4 off */
int oracle_qp_solve_ld(int n, int m, const double *C, const double *dvec, const double *A, const double *b,
                       double *x_out, double *u_out, int *iterations)
{
  return qp_solve_t<long double>(n, m, C, dvec, A, b, x_out, u_out, iterations);
}

} /* extern "C" */
