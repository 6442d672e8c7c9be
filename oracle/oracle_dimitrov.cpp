/* oracle/oracle_dimitrov.cpp - TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain restatement of the Dimitrov2008 path around the PLDP solver:
 *   ComputeConvexHull::DoComputeConvexHull                       src/Mathematics/ConvexHull.cpp:87-203
 *   FootConstraintsAsLinearSystem::FindSimilarConstraints        src/Mathematics/FootConstraintsAsLinearSystem.cpp:53-92
 *   FootConstraintsAsLinearSystem::ComputeLinearSystem           :94-256
 *   FootConstraintsAsLinearSystem::BuildLinearConstraintInequalities   :258-539
 *   ZMPConstrainedQPFastFormulation::InitializeMatrixPbConstants src/ZMPRefTrajectoryGeneration/ZMPConstrainedQPFastFormulation.cpp:158-246
 *   ...::BuildingConstantPartOfTheObjectiveFunction(+QLDANDLQ)   :390-614   (including `lterm2 = alpha * VPu^T`, sic, :525-533)
 *   ...::BuildingConstantPartOfConstraintMatrices                :616-720
 *   ...::BuildConstraintMatrices                                 :759-1022
 *   ...::BuildZMPTrajectoryFromFootTrajectory (PLDP branch)      :1095-1480
 *   LinearizedInvertedPendulum2D::Interpolation / OneIteration   src/PreviewControl/LinearizedInvertedPendulum2D.cpp:157-264
 *
 * Pins: the convex hull and BuildLinearConstraintInequalities are checked against the reference's OWN object code
 * (ConvexHull.cpp, FootConstraintsAsLinearSystem.cpp compiled where they lie into oracle/_ref, recipe oracle/Makefile)
 * on the feet trajectories of the four TestKajita2003 profiles and on random point sets (tests/test_dimitrov.py);
 * the closed loop runs the PLDP restatement (oracle_pldp.cpp, bitwise equal to the reference's PLDPSolver object
 * code) and, in the tests, the reference's PLDPSolver object itself.  ZMPConstrainedQPFastFormulation.cpp does not
 * compile here (uBLAS/jrl-mal/LAPACK MAL_INVERSE) and the reference ships no test or datref for this generator
 * (the PGI branch that selects it is commented out): the loop itself is PARITY UNPINNED by golden vectors.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/walkgen_b200.h"

extern "C" int oracle_pldp_solve_sim(int N, const double *iPu, const double *Px, const double *Pu, const double *D, int m,
                                     const double *A, const double *b, const double *ZMPRef, const double *XkYk, double *X,
                                     int n_removed, int starting, void *hot, int hot_start, int max_iter, int *info,
                                     int *active_out, const int *similar);
extern "C" int oracle_pldp_solve(int N, const double *iPu, const double *Px, const double *Pu, const double *D, int m,
                                 const double *DPu, const double *DPx, const double *ZMPRef, const double *XkYk,
                                 double *X, int n_removed, int starting, void *hot, int hot_start, int max_iter,
                                 int *info, int *active);
extern "C" void oracle_optcholesky_full(int n, const double *A, double *L);
extern "C" void oracle_optcholesky_inverse(int n, int size, const double *L, double *iL);

namespace {

struct Pt { double col, row; };

double cross0(const Pt &p0, const Pt &s1, const Pt &s2)
{
  const double x1 = s1.col - p0.col, x2 = s2.col - p0.col, y1 = s1.row - p0.row, y2 = s2.row - p0.row;
  return x1 * y2 - x2 * y1;
}

/* ConvexHull.cpp:87-203.  The reference keeps the candidates in a std::set ordered by "cross product about p0 > 0";
 * here: a vector kept in that order (same result whenever the comparison is a strict weak order on the candidates,
 * which it is once points of equal polar angle have been merged as the reference does). */
void convex_hull(const std::vector<Pt> &in, std::vector<Pt> &hull)
{
  if (in.empty()) return;
  Pt p0 = in[0];
  for (size_t i = 0; i < in.size(); ++i)
    if (in[i].row < p0.row) p0 = in[i];
  std::vector<Pt> lst;
  for (size_t i = 0; i < in.size(); ++i) {
    bool ins = true;
    for (size_t k = 0; k < lst.size();) {
      bool del = false;
      if (cross0(p0, lst[k], in[i]) == 0.0) {
        const double x1 = lst[k].col - p0.col, y1 = lst[k].row - p0.row, x2 = in[i].col - p0.col, y2 = in[i].row - p0.row;
        const double d1 = sqrt(x1 * x1 + y1 * y1), d2 = sqrt(x2 * x2 + y2 * y2);
        if (d1 <= d2) del = true; else ins = false;
      }
      if (del) lst.erase(lst.begin() + k); else ++k;
    }
    if (ins) {
      size_t pos = 0;
      while (pos < lst.size() && cross0(p0, lst[pos], in[i]) > 0.0) ++pos;   /* after every element that sorts before it */
      lst.insert(lst.begin() + pos, in[i]);
    }
  }
  hull.push_back(p0);
  if (lst.size() < 2) { for (const Pt &p : lst) hull.push_back(p); return; }   /* the reference dereferences end() here */
  hull.push_back(lst[0]);
  hull.push_back(lst[1]);
  for (size_t k = 2; k < lst.size(); ++k) {
    const Pt pi = lst[k];
    bool ok;
    do {
      if (hull.size() >= 2) {
        const Pt s1 = hull[hull.size() - 1], s2 = hull[hull.size() - 2];
        const double x1 = s1.col - s2.col, x2 = pi.col - s2.col, y1 = s1.row - s2.row, y2 = pi.row - s2.row;
        ok = (x1 * y2 - x2 * y1) > 0.0;
      } else ok = true;
      if (!ok) hull.pop_back();
    } while (!ok);
    hull.push_back(pi);
  }
}

/* one edge of ComputeLinearSystem (:151-193 for edges i -> i+1, :207-243 for the closing edge n-1 -> 0; the two copies
 * differ only in which end point supplies the intercept: `from` in the loop, `to` (= point 0) in the closing edge) */
void edge(const Pt &from, const Pt &to, const Pt &icpt, double &a, double &b, double &c)
{
  if (fabs(to.col - from.col) > 1e-7) {
    double y1, x1, y2, x2, lmul = -1.0;
    if (to.col < from.col) { lmul = 1.0; y2 = from.row; y1 = to.row; x2 = from.col; x1 = to.col; }
    else { y2 = to.row; y1 = from.row; x2 = to.col; x1 = from.col; }
    a = (y2 - y1) / (x2 - x1);
    b = (icpt.row - a * icpt.col);
    a = lmul * a; b = lmul * b; c = -lmul;
  } else {
    c = 0.0; a = -1.0; b = to.col;
    if (to.row < from.row) { a = -a; b = -b; }
  }
}

int linear_system(const std::vector<Pt> &v, wg_lci &o, int merge)
{
  const unsigned n = (unsigned)v.size();
  double C0 = 0.0, C1 = 0.0;
  for (unsigned i = 0; i + 1 < n; ++i) {
    C0 += v[i].col; C1 += v[i].row;
    double a, b, c;
    edge(v[i], v[i + 1], v[i], a, b, c);
    o.A[i][0] = a; o.A[i][1] = c; o.B[i] = b;
  }
  C0 += v[n - 1].col; C1 += v[n - 1].row;
  C0 /= (double)n; C1 /= (double)n;
  double a, b, c;
  edge(v[n - 1], v[0], v[0], a, b, c);
  o.A[n - 1][0] = a; o.A[n - 1][1] = c; o.B[n - 1] = b;
  o.center[0] = C0; o.center[1] = C1;
  o.rows = (int)n;
  if (merge) {
    /* NOT in the reference (wg_dimitrov_params.merge_duplicate_rows): drop a half-plane that repeats its predecessor */
    auto near = [&](int i, int j) {
      return fabs(o.A[i][0] - o.A[j][0]) <= 1e-9 && fabs(o.A[i][1] - o.A[j][1]) <= 1e-9 && fabs(o.B[i] - o.B[j]) <= 1e-9;
    };
    int k = 0;
    for (int i = 0; i < (int)n; ++i) {
      if (k > 0 && near(i, k - 1)) continue;
      o.A[k][0] = o.A[i][0]; o.A[k][1] = o.A[i][1]; o.B[k] = o.B[i];
      ++k;
    }
    if (k > 1 && near(k - 1, 0)) --k;
    for (int i = k; i < (int)n; ++i) { o.A[i][0] = 0.0; o.A[i][1] = 0.0; o.B[i] = 0.0; }
    o.rows = k;
  }
  /* W = A C (rows 0 and 1 only are tested, :244-249) */
  const double W0 = (o.A[0][0] * C0 + o.A[0][1] * C1) + o.B[0], W1 = (o.A[1][0] * C0 + o.A[1][1] * C1) + o.B[1];
  return (W0 < 0 || W1 < 0) ? -1 : 0;
}

void find_similar(wg_lci &o)
{
  for (int i = 0; i < WG_LCI_MAX_ROWS; ++i) o.similar[i] = 0;
  const int n = o.rows;
  if (n == 4) {
    if (o.A[0][0] == -o.A[2][0] && o.A[0][1] == -o.A[2][1]) o.similar[2] = -2;
    if (o.A[1][0] == -o.A[3][0] && o.A[1][1] == -o.A[3][1]) o.similar[3] = -2;
  } else if (n == 6) {
    for (int k = 0; k < 3; ++k)
      if (o.A[k][0] == -o.A[k + 3][0] && o.A[k][1] == -o.A[k + 3][1]) o.similar[k + 3] = -3;
  }
}

const double lxcoefs[4] = {1.0, 1.0, -1.0, -1.0};
const double lycoefs[4] = {-1.0, 1.0, 1.0, -1.0};

void foot_corners(const double *f /* x y z theta */, double hw, double hh, Pt *out)
{
  const double lx = f[0], ly = f[1];
  const double s_t = sin(f[3] * M_PI / 180.0), c_t = cos(f[3] * M_PI / 180.0);
  for (unsigned j = 0; j < 4; ++j) {
    out[j].col = lx + (lxcoefs[j] * hw * c_t - lycoefs[j] * hh * s_t);
    out[j].row = ly + (lxcoefs[j] * hw * s_t + lycoefs[j] * hh * c_t);
  }
}

struct DimConsts {
  int N;
  double T, Ts, zc;
  std::vector<double> iPu, Px, Pu, iLQ, OptB, OptC;   /* one-axis blocks, row-major */
};

/* plain Gauss-Jordan with partial pivoting: stands in for MAL_INVERSE (LAPACK dgetrf/dgetri in jrl-mal; parity
 * unpinned, SURVEY 8c "third-party arithmetic") */
void invert(int n, const double *A, double *inv)
{
  std::vector<double> M(A, A + n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) inv[i * n + j] = (i == j);
  for (int c = 0; c < n; ++c) {
    int p = c;
    for (int r = c + 1; r < n; ++r) if (fabs(M[r * n + c]) > fabs(M[p * n + c])) p = r;
    if (p != c) for (int j = 0; j < n; ++j) { std::swap(M[p * n + j], M[c * n + j]); std::swap(inv[p * n + j], inv[c * n + j]); }
    const double d = M[c * n + c];
    for (int j = 0; j < n; ++j) { M[c * n + j] /= d; inv[c * n + j] /= d; }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = M[r * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < n; ++j) { M[r * n + j] -= f * M[c * n + j]; inv[r * n + j] -= f * inv[c * n + j]; }
    }
  }
}

void make_consts(int N, double T, double zc, double alpha, double beta, DimConsts &K)
{
  K.N = N; K.T = T; K.zc = zc;
  std::vector<double> PPu(N * N, 0.0), VPu(N * N, 0.0), PPx(N * 3), VPx(N * 3);
  for (int i = 0; i < N; ++i) {
    VPx[i * 3 + 0] = 0.0; VPx[i * 3 + 1] = 1.0; VPx[i * 3 + 2] = (i + 1) * T;
    PPx[i * 3 + 0] = 1.0; PPx[i * 3 + 1] = (i + 1) * T; PPx[i * 3 + 2] = (i + 1) * (i + 1) * T * T * 0.5;
    for (int j = 0; j <= i; ++j) {
      VPu[i * N + j] = (2 * (i - j) + 1) * T * T * 0.5;
      PPu[i * N + j] = (1 + 3 * (i - j) + 3 * (i - j) * (i - j)) * T * T * T / 6.0;
    }
  }
  /* OptA (upper-left block) = I + beta PPu^T PPu + alpha VPu^T  (sic) */
  std::vector<double> Q(N * N), LQ(N * N, 0.0), iLQ(N * N, 0.0);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double t = 0.0;
      for (int k = 0; k < N; ++k) t += PPu[k * N + i] * PPu[k * N + j];
      Q[i * N + j] = ((i == j ? 1.0 : 0.0) + beta * t) + alpha * VPu[j * N + i];
    }
  oracle_optcholesky_full(N, Q.data(), LQ.data());
  oracle_optcholesky_inverse(N, N, LQ.data(), iLQ.data());
  K.iLQ = iLQ;
  /* OptB = iLQ (alpha VPu^T VPx + beta PPu^T PPx), OptC = iLQ (beta PPu^T) */
  std::vector<double> B0(N * 3), C0(N * N);
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < 3; ++j) {
      double tv = 0.0, tp = 0.0;
      for (int k = 0; k < N; ++k) { tv += VPu[k * N + i] * VPx[k * 3 + j]; tp += PPu[k * N + i] * PPx[k * 3 + j]; }
      B0[i * 3 + j] = alpha * tv + beta * tp;
    }
    for (int j = 0; j < N; ++j) C0[i * N + j] = beta * PPu[j * N + i];
  }
  K.OptB.assign(N * 3, 0.0); K.OptC.assign(N * N, 0.0);
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < 3; ++j) { double t = 0.0; for (int k = 0; k < N; ++k) t += iLQ[i * N + k] * B0[k * 3 + j]; K.OptB[i * 3 + j] = t; }
    for (int j = 0; j < N; ++j) { double t = 0.0; for (int k = 0; k < N; ++k) t += iLQ[i * N + k] * C0[k * N + j]; K.OptC[i * N + j] = t; }
  }
  /* Pu' (ptPu[k*N+i], k <= i), m_Pu = iLQ Pu', iPu = inverse */
  std::vector<double> PuT(N * N, 0.0);
  for (int i = 0; i < N; ++i)
    for (int k = 0; k <= i; ++k)
      PuT[k * N + i] = ((1 + 3 * (i - k) + 3 * (i - k) * (i - k)) * T * T * T / 6.0 - T * zc / 9.81);
  K.Pu.assign(N * N, 0.0);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      double t = 0.0;
      for (int k = 0; k < N; ++k) t += iLQ[i * N + k] * PuT[k * N + j];
      K.Pu[i * N + j] = t;
    }
  K.iPu.assign(N * N, 0.0);
  invert(N, K.Pu.data(), K.iPu.data());
  K.Px.resize(N * 3);
  for (int li = 0; li < N; ++li) {
    K.Px[li * 3 + 0] = 1.0;
    K.Px[li * 3 + 1] = (double)(1.0 + li) * T;
    K.Px[li * 3 + 2] = (li + 1.0) * (li + 1.0) * T * T * 0.5 - zc / 9.81;
  }
}

}  // namespace

extern "C" {

int oracle_convex_hull(int n, const double *xy, double *hull_xy, int cap)
{
  std::vector<Pt> in(n), out;
  for (int i = 0; i < n; ++i) { in[i].col = xy[2 * i]; in[i].row = xy[2 * i + 1]; }
  convex_hull(in, out);
  for (int i = 0; i < (int)out.size() && i < cap; ++i) { hull_xy[2 * i] = out[i].col; hull_xy[2 * i + 1] = out[i].row; }
  return (int)out.size();
}

/* BuildLinearConstraintInequalities.  feet [n][4] = x, y, z, theta(deg); step_type / time: the left foot's.
 * Returns the number of polygons found (the first `cap` are stored). */
int oracle_fcals_build(int n, const double *left, const double *right, const int *step_type, const double *time,
                       double sole_length, double sole_width, double cx, double cy, int cap, wg_lci *out, int merge)
{
  double lhw = sole_length, lhh = sole_width, rhw = sole_length, rhh = sole_width;
  rhw *= 0.5; rhh *= 0.5; lhw *= 0.5; lhh *= 0.5;
  lhh -= cy; rhh -= cy;
  lhw -= cx; rhw -= cx;
  int State = 0, np = 0;
  for (int i = 0; i < n; ++i) {
    const double *L = left + 4 * i, *R = right + 4 * i;
    int ComputeCH = 0;
    if (i == 0) { ComputeCH = 1; State = 3; }
    if (step_type[i] >= 10) {
      if (State != 3) ComputeCH = 1;
      State = 3;
    } else {
      const double thr = 0.00001;
      if (L[2] > thr) { if (State != 2) ComputeCH = 1; State = 2; }
      else if (R[2] > thr) { if (State != 1) ComputeCH = 1; State = 1; }
      else if (R[2] < thr && L[2] < thr) { if (State != 3) ComputeCH = 1; State = 3; }
    }
    if (ComputeCH) {
      std::vector<Pt> hull;
      if (State == 3) {
        std::vector<Pt> pts(8);
        foot_corners(L, lhw, lhh, pts.data());
        foot_corners(R, rhw, rhh, pts.data() + 4);
        convex_hull(pts, hull);
      } else {
        hull.resize(4);
        if (L[2] < R[2]) foot_corners(L, lhw, lhh, hull.data());
        else foot_corners(R, rhw, rhh, hull.data());
      }
      wg_lci o;
      std::memset(&o, 0, sizeof o);
      if ((int)hull.size() > WG_LCI_MAX_ROWS) return -2;
      o.rc = linear_system(hull, o, merge);
      find_similar(o);
      o.t_start = time[i];
      o.first_sample = i;
      o.state = State;
      if (np > 0 && np - 1 < cap) out[np - 1].t_end = time[i];
      if (np < cap) out[np] = o;
      ++np;
    }
    if (i == n - 1 && np > 0 && np - 1 < cap) out[np - 1].t_end = time[i];
  }
  return np;
}

/* InitConstants for one axis block.  All outputs row-major: iPu, Pu, iLQ, OptC [N][N]; Px, OptB [N][3]. */
void oracle_dimitrov_constants(int N, double T, double zc, double alpha, double beta, double *iPu, double *Px, double *Pu,
                               double *iLQ, double *OptB, double *OptC)
{
  DimConsts K;
  make_consts(N, T, zc, alpha, beta, K);
  if (iPu) std::memcpy(iPu, K.iPu.data(), sizeof(double) * N * N);
  if (Px) std::memcpy(Px, K.Px.data(), sizeof(double) * N * 3);
  if (Pu) std::memcpy(Pu, K.Pu.data(), sizeof(double) * N * N);
  if (iLQ) std::memcpy(iLQ, K.iLQ.data(), sizeof(double) * N * N);
  if (OptB) std::memcpy(OptB, K.OptB.data(), sizeof(double) * N * 3);
  if (OptC) std::memcpy(OptC, K.OptC.data(), sizeof(double) * N * N);
}

/* BuildConstraintMatrices (:759-1022) + the D vector (:1268-1276) for one StartingTime.  consts: the arrays of
 * oracle_dimitrov_constants (N = 16).  DPu is (m+1) x 2N column-major, zero-initialised; returns m (< 0: -1 no polygon
 * covers StartingTime, -3 ran past the last polygon).  first[0] = index of the polygon found, first[1] = its rows. */
int oracle_dimitrov_build_constraints(int N, double T, double StartingTime, int np, const wg_lci *lci, const double *Px,
                                      const double *Pu, const double *OptB, const double *OptC, const double *xk,
                                      double *DPu, double *DPx, double *ZMPRef, double *D, int *first)
{
  int it = 0;
  while (it < np) {
    if (StartingTime >= lci[it].t_start && StartingTime <= lci[it].t_end) break;
    ++it;
  }
  if (it == np) return -1;
  const int store = it;
  unsigned m = 0;
  for (int i = 0; i < N; ++i) {
    const double ltime = StartingTime + i * T;
    if (ltime > lci[it].t_end) ++it;
    if (it == np) break;
    m += lci[it].rows;
  }
  if (it == np) return -3;
  it = store;
  first[0] = store; first[1] = lci[store].rows;
  std::memset(DPu, 0, sizeof(double) * (m + 1) * 2 * N);
  unsigned r = 0;
  for (int i = 0; i < N; ++i) {
    const double ltime = StartingTime + i * T;
    if (ltime > lci[it].t_end) ++it;
    ZMPRef[i] = lci[it].center[0];
    ZMPRef[i + N] = lci[it].center[1];
    for (int j = 0; j < lci[it].rows; ++j) {
      DPx[r] = (xk[0] * Px[i * 3 + 0] + xk[1] * Px[i * 3 + 1] + xk[2] * Px[i * 3 + 2]) * lci[it].A[j][0] +
               (xk[3] * Px[i * 3 + 0] + xk[4] * Px[i * 3 + 1] + xk[5] * Px[i * 3 + 2]) * lci[it].A[j][1] + lci[it].B[j];
      for (int k = 0; k < N; ++k) {
        DPu[r + k * (m + 1)] = lci[it].A[j][0] * Pu[k * N + i];
        DPu[r + (k + N) * (m + 1)] = lci[it].A[j][1] * Pu[k * N + i];
      }
      ++r;
    }
  }
  /* OptD = OptB xk - OptC ZMPRef, row by row (block diagonal: the other axis contributes exact zeros) */
  for (int ax = 0; ax < 2; ++ax)
    for (int i = 0; i < N; ++i) {
      double t1 = 0.0, t2 = 0.0;
      for (int j = 0; j < N; ++j) t1 += OptC[i * N + j] * ZMPRef[j + N * ax];
      for (int j = 0; j < 3; ++j) t2 += OptB[i * 3 + j] * xk[3 * ax + j];
      D[i + N * ax] = t2 - t1;
    }
  return (int)m;
}

/* m_SimilarConstraints as BuildConstraintMatrices fills it (:889): the SimilarConstraints of the polygon of each previewed
 * sample, row by row, for the same polygon walk as oracle_dimitrov_build_constraints.  Returns m. */
int oracle_dimitrov_similar(int N, double T, double StartingTime, int np, const wg_lci *lci, int *similar)
{
  int it = 0;
  while (it < np) {
    if (StartingTime >= lci[it].t_start && StartingTime <= lci[it].t_end) break;
    ++it;
  }
  if (it == np) return -1;
  int r = 0;
  for (int i = 0; i < N; ++i) {
    const double ltime = StartingTime + i * T;
    if (ltime > lci[it].t_end) ++it;
    if (it == np) return -3;
    for (int j = 0; j < lci[it].rows; ++j) similar[r++] = lci[it].similar[j];
  }
  return r;
}

/* number of periods of the loop `for (StartingTime = 0; StartingTime < EndingTime - N*T; StartingTime += T)` for a
 * feet buffer of n samples whose clock is the sampling period accumulated sample by sample */
long oracle_dimitrov_period_count(int N, double T, double Ts, long n)
{
  double t = 0.0;
  for (long i = 1; i < n; ++i) t += Ts;
  long c = 0;
  for (double st = 0.0; st < t - N * T; st += T) ++c;
  return c;
}

/* The whole generator after ZMPDiscretization: feet buffers -> CoM / ZMP at 5 ms.
 *   left/right [n][4], step_type [n]; par = wg_dimitrov_params
 *   com [n][6], zmp [n][2] (rows the loop reaches are overwritten), periods [cap_periods]
 * Returns the number of periods (< 0 on failure: -(100+k) the solve of period k failed - IFAIL or the exit(0) path of
 * the reference -, -2 no polygon, -4 polygon capacity). */
long oracle_dimitrov_run(const wg_dimitrov_params *par, long n, const double *left, const double *right,
                         const int *step_type, double *com, double *zmp, long cap_periods, wg_dimitrov_period *periods,
                         int use_hot_start)
{
  const int N = 16;
  DimConsts K;
  make_consts(N, par->T, par->com_height, par->alpha, par->beta, K);
  std::vector<double> time(n);
  { double t = 0.0; for (long i = 0; i < n; ++i) { time[i] = t; t += par->sampling_period; } }
  std::vector<wg_lci> lci(4096);
  const int np = oracle_fcals_build((int)n, left, right, step_type, time.data(), par->sole_length, par->sole_width,
                                    par->constraint_x, par->constraint_y, (int)lci.size(), lci.data(), par->merge_duplicate_rows);
  if (np < 0 || np > (int)lci.size()) return -4;
  const double T = par->T, Ts = par->sampling_period;
  const int interval = (int)(T / Ts);
  double cx[3] = {0, 0, 0}, cy[3] = {0, 0, 0};
  std::vector<double> DPu((8 * N + 1) * 2 * N), DPx(8 * N + 1), ZMPRef(2 * N), D(2 * N), X(2 * N);
  struct { double prev_zmp[32]; int prev_active[32]; int n_prev; int pad_; } hot;
  std::memset(&hot, 0, sizeof hot);
  bool starting = true;
  unsigned removed = 0;
  long li = 0;
  const int max_iter = par->max_iterations > 0 ? par->max_iterations : 128;
  for (double ST = 0.0; ST < lci[np - 1].t_end - N * T; ST += T, ++li) {
    const double xk[6] = {cx[0], cx[1], cx[2], cy[0], cy[1], cy[2]};
    int first[2];
    const int m = oracle_dimitrov_build_constraints(N, T, ST, np, lci.data(), K.Px.data(), K.Pu.data(), K.OptB.data(),
                                                    K.OptC.data(), xk, DPu.data(), DPx.data(), ZMPRef.data(), D.data(), first);
    if (m < 0) return -2;
    int info[4], act[32];
    /* m_SimilarConstraints is handed to the solver as the reference does (:1336); the reuse it enables is bit-neutral
     * for these flags (the flagged row is the exact negation of the row it points to) */
    int similar[8 * 16];
    oracle_dimitrov_similar(N, T, ST, np, lci.data(), similar);
    int rc = oracle_pldp_solve_sim(N, K.iPu.data(), K.Px.data(), K.Pu.data(), D.data(), m, DPu.data(), DPx.data(),
                                   ZMPRef.data(), xk, X.data(), (int)removed, starting ? 1 : 0, &hot, use_hot_start,
                                   max_iter, info, act, similar);
    if ((info[1] == 1 || info[1] == 2) && par->cold_restart) {
      /* the reference prints "PB ON constraint" and calls exit(0) here; cold_restart solves the period again from the
       * cold start point (StartingSequence = true, no kept constraints) */
      hot.n_prev = 0;
      rc = oracle_pldp_solve_sim(N, K.iPu.data(), K.Px.data(), K.Pu.data(), D.data(), m, DPu.data(), DPx.data(),
                                 ZMPRef.data(), xk, X.data(), 0, 1, &hot, use_hot_start, max_iter, info, act, similar);
      if (info[1] == 0) info[1] = 5;
    }
    starting = false;
    removed = (unsigned)first[1];
    /* NewX = iLQ^T X (only entries 0 and N are used) */
    double jx = 0.0, jy = 0.0;
    for (int j = 0; j < N; ++j) jx += K.iLQ[j * N + 0] * X[j];
    for (int j = 0; j < N; ++j) jy += K.iLQ[j * N + 0] * X[j + N];
    if (periods && li < cap_periods) {
      wg_dimitrov_period &P = periods[li];
      std::memset(&P, 0, sizeof P);
      P.t_start = ST;
      std::memcpy(P.xk, xk, sizeof xk);
      P.jerk_x = jx; P.jerk_y = jy;
      P.m = m; P.n_first = first[1];
      P.rc = info[0]; P.status = info[1]; P.iterations = info[2]; P.n_active = info[3];
      for (int k = 0; k < 32; ++k) P.active[k] = act[k];
    }
    if (rc != 0 || (info[1] != 0 && info[1] != 5)) return -(100 + li);   /* IFAIL / exit(0) in the reference */
    /* Interpolation(COMStates, ZMPRefPositions, li*interval, jx, jy): interval+1 samples */
    const long cur = li * interval;
    const int loopEnd = (int)std::min<long>(interval, n - 1 - cur);
    for (int lk = 0; lk <= loopEnd; ++lk) {
      const double s = (lk + 1) * Ts;
      double *c = com + 6 * (cur + lk);
      c[0] = cx[0] + s * cx[1] + 0.5 * s * s * cx[2] + s * s * s * jx / 6.0;
      c[1] = cx[1] + s * cx[2] + 0.5 * s * s * jx;
      c[2] = cx[2] + s * jx;
      c[3] = cy[0] + s * cy[1] + 0.5 * s * s * cy[2] + s * s * s * jy / 6.0;
      c[4] = cy[1] + s * cy[2] + 0.5 * s * s * jy;
      c[5] = cy[2] + s * jy;
      const double C2 = -par->com_height / 9.81;
      zmp[2 * (cur + lk)] = 1.0 * c[0] + 0.0 * c[1] + C2 * c[2];
      zmp[2 * (cur + lk) + 1] = 1.0 * c[3] + 0.0 * c[4] + C2 * c[5];
    }
    /* OneIteration: x = A x + B u */
    const double A01 = T, A02 = T * T / 2.0, A12 = T, B0 = T * T * T / 6.0, B1 = T * T / 2.0, B2 = T;
    double nx[3], ny[3];
    nx[0] = ((1.0 * cx[0] + A01 * cx[1]) + A02 * cx[2]) + jx * B0;
    nx[1] = ((0.0 * cx[0] + 1.0 * cx[1]) + A12 * cx[2]) + jx * B1;
    nx[2] = ((0.0 * cx[0] + 0.0 * cx[1]) + 1.0 * cx[2]) + jx * B2;
    ny[0] = ((1.0 * cy[0] + A01 * cy[1]) + A02 * cy[2]) + jy * B0;
    ny[1] = ((0.0 * cy[0] + 1.0 * cy[1]) + A12 * cy[2]) + jy * B1;
    ny[2] = ((0.0 * cy[0] + 0.0 * cy[1]) + 1.0 * cy[2]) + jy * B2;
    for (int k = 0; k < 3; ++k) { cx[k] = nx[k]; cy[k] = ny[k]; }
  }
  return li;
}

}  /* extern "C" */
