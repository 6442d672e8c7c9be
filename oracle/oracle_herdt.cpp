/* oracle/oracle_herdt.cpp - TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Restatement of the reference's Herdt2010 velocity-referenced MPC, function by function:
 *   ZMPVelocityReferencedQP ctor / InitOnLine / OnLine   src/ZMPRefTrajectoryGeneration/ZMPVelocityReferencedQP.cpp:56-135, 213-319, 324-458
 *   SupportFSM::update_vel_reference / set_support_state  src/PreviewControl/SupportFSM.cpp:58-153
 *   GeneratorVelRef::preview_support_states               src/ZMPRefTrajectoryGeneration/generator-vel-ref.cpp:71-134
 *   GeneratorVelRef::generate_selection_matrices          generator-vel-ref.cpp:138-208
 *   GeneratorVelRef::compute_global_reference             generator-vel-ref.cpp:212-229
 *   GeneratorVelRef::build_invariant_part/update_problem  generator-vel-ref.cpp:588-674
 *   GeneratorVelRef::build_inequalities_xxx, build_constraints_xxx   generator-vel-ref.cpp:285-474, 555-584
 *   RigidBodySystem::compute_dyn_cjerk                    src/PreviewControl/rigid-body-system.cpp:377-452
 *   RelativeFeetInequalities (hulls, half planes)         src/Mathematics/relative-feet-inequalities.cpp:40-319
 *   FootHalfSize                                          src/Mathematics/FootHalfSize.cpp:52-100
 *   QPProblem (dummy row, QLD calling convention)         src/ZMPRefTrajectoryGeneration/qp-problem.cpp:246-293, 411-547
 *   LinearizedInvertedPendulum2D::Interpolation/OneIter.  src/PreviewControl/LinearizedInvertedPendulum2D.cpp:157-264
 *   OrientationsPreview                                   src/ZMPRefTrajectoryGeneration/OrientationsPreview.cpp:40-431
 *   OnLineFootTrajectoryGeneration                        src/FootTrajectoryGeneration/OnLineFootTrajectoryGeneration.cpp:51-346
 *   Polynome3/4/5                                         src/Mathematics/PolynomeFoot.cpp:33-240, Polynome.cpp:44-75
 *   CoMAndFootOnlyStrategy::OneGlobalStepOfControl        src/GlobalStrategyManagers/CoMAndFootOnlyStrategy.cpp:56-124
 *   PGI::RunOneStepOfTheControlLoop (Herdt branch)        src/PatternGeneratorInterfacePrivate.cpp:1246-1336
 *
 * The QP itself is solved by the reference's own ql0001_ (oracle/_ref/libwalkgen_ref.so, loaded with
 * dlopen when present) or, failing that, by oracle_qp_solve (oracle/oracle_qp.cpp).
 *
 * Robot data that the reference reads from jrl-dynamics' sample robot (not in the container) are
 * parameters here: sole size (fitted from the datref: 0.25 x 0.14 m), hip-yaw limits and
 * velocity bound (OrientationsPreview.cpp:48-68; unknown -> rotation segments are unpinned).
 */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <vector>
#include <algorithm>
#include <dlfcn.h>
#include "../include/walkgen_b200.h"

extern "C" int oracle_qp_solve(int n, int m, const double *C, const double *d, const double *A, const double *b,
                               double *x, double *u, int *iterations);

namespace {

const int N = WG_HERDT_N;
enum { LEFT = 0, RIGHT = 1 };
enum { SS = 0, DS = 1 };

struct Support {  /* support_state_t, privatepgtypes.hh:291-320 */
  int Phase = DS, Foot = LEFT;
  unsigned NbStepsLeft = 0, StepNumber = 0, NbInstants = 0;
  double TimeLimit = 0, StartTime = 0, X = 0, Y = 0, Yaw = 0;
  bool StateChanged = false;
  /* operator= of the reference does not copy NbInstants (privatepgtypes.cpp:33-49) */
  void assign_like_reference(const Support &o)
  {
    unsigned keep = NbInstants;
    *this = o;
    NbInstants = keep;
  }
};

struct Foot {  /* FootAbsolutePosition, pgtypes.hh:141-170 */
  double x = 0, y = 0, z = 0, theta = 0, omega = 0, omega2 = 0;
  double dx = 0, dy = 0, dz = 0, dtheta = 0, domega = 0, domega2 = 0;
  double ddx = 0, ddy = 0, ddz = 0, ddtheta = 0, ddomega = 0, ddomega2 = 0;
  double time = 0;
  int stepType = 0;
};
struct Com {  /* COMState */
  double x[3] = {0, 0, 0}, y[3] = {0, 0, 0}, z[3] = {0, 0, 0}, yaw[3] = {0, 0, 0};
};
struct Zmp {
  double px = 0, py = 0, pz = 0, theta = 0, time = 0;
  int stepType = 0;
};

struct Poly {  /* Polynome, Polynome.cpp:44-75 */
  std::vector<double> c;
  explicit Poly(int deg) : c(deg + 1, 0.0) {}
  double val(double t) const { double r = 0, pt = 1; for (size_t i = 0; i < c.size(); ++i) { r += c[i] * pt; pt *= t; } return r; }
  double d1(double t) const { double r = 0, pt = 1; for (size_t i = 1; i < c.size(); ++i) { r += i * c[i] * pt; pt *= t; } return r; }
  double d2(double t) const { double r = 0, pt = 1; for (size_t i = 2; i < c.size(); ++i) { r += i * (i - 1) * c[i] * pt; pt *= t; } return r; }
  int degree() const { return (int)c.size() - 1; }
};
/* Polynome3::SetParametersWithInitPosInitSpeed, PolynomeFoot.cpp:57-78 */
void poly3_init(Poly &p, double FT, double FP, double ip, double is)
{
  p.c[0] = ip; p.c[1] = is;
  double tmp = FT * FT;
  if (FT == 0.0) { p.c[2] = 0; p.c[3] = 0; }
  else { p.c[2] = (3 * FP - 3 * ip - 2 * is * FT) / tmp; p.c[3] = (is * FT + 2 * ip - 2 * FP) / (tmp * FT); }
}
/* Polynome4::SetParameters, PolynomeFoot.cpp:100-121 */
void poly4_set(Poly &p, double FT, double MP)
{
  p.c[0] = 0; p.c[1] = 0;
  double tmp = FT * FT;
  if (MP == 0.0 || tmp == 0.0) { p.c[2] = p.c[3] = p.c[4] = 0; }
  else { p.c[2] = 16.0 * MP / tmp; tmp *= FT; p.c[3] = -32.0 * MP / tmp; tmp *= FT; p.c[4] = 16.0 * MP / tmp; }
}
/* Polynome5::SetParameters(FT,FP,InitPos,InitSpeed,InitAcc), PolynomeFoot.cpp:226-240 */
void poly5_set(Poly &p, double FT, double FP, double ip, double is, double ia)
{
  p.c[0] = ip; p.c[1] = is; p.c[2] = ia / 2.0;
  double tmp = FT * FT * FT;
  p.c[3] = (-3.0 / 2.0 * ia * FT * FT - 6.0 * is * FT - 10.0 * ip + 10.0 * FP) / tmp;
  tmp *= FT;
  p.c[4] = (3.0 / 2.0 * ia * FT * FT + 8.0 * is * FT + 15.0 * ip - 15.0 * FP) / tmp;
  tmp *= FT;
  p.c[5] = (-1.0 / 2.0 * ia * FT * FT - 3.0 * is * FT - 6.0 * ip + 6.0 * FP) / tmp;
}

struct Dyn { double S[N][3]; double U[N][N]; };

/* RigidBodySystem::compute_dyn_cjerk, rigid-body-system.cpp:377-452 */
void dyn_velocity(Dyn &D, double T)
{
  for (int i = 0; i < N; ++i) {
    D.S[i][0] = 0.0; D.S[i][1] = 1.0; D.S[i][2] = (i + 1) * T;
    for (int j = 0; j < N; ++j) D.U[i][j] = (j <= i) ? (2 * (i - j) + 1) * T * T * 0.5 : 0.0;
  }
}
void dyn_cop(Dyn &D, double T, double h)
{
  for (int i = 0; i < N; ++i) {
    D.S[i][0] = 1.0; D.S[i][1] = (i + 1) * T; D.S[i][2] = (i + 1) * (i + 1) * T * T * 0.5 - h / 9.81;
    for (int j = 0; j < N; ++j)
      D.U[i][j] = (j <= i) ? (1 + 3 * (i - j) + 3 * (i - j) * (i - j)) * T * T * T / 6.0 - T * h / 9.81 : 0.0;
  }
}

typedef int (*ql0001_fn)(int *, int *, int *, int *, int *, int *, double *, double *, double *, double *, double *,
                         double *, double *, double *, int *, int *, int *, double *, int *, int *, int *, double *);
ql0001_fn g_ql = nullptr;
bool g_ql_tried = false;
char g_ref_path[1024] = "";

ql0001_fn load_ql()
{
  if (!g_ql_tried) {
    g_ql_tried = true;
    if (g_ref_path[0]) {
      void *h = dlopen(g_ref_path, RTLD_NOW | RTLD_LOCAL);
      if (h) g_ql = (ql0001_fn)dlsym(h, "ref_ql0001");
    }
  }
  return g_ql;
}

/* Dense problem in the exact layout QPProblem::solve hands to ql0001_ (qp-problem.cpp:246-280):
 * Q n x n column-major; DU (m_+1) x n column-major, row 0 all-zero dummy; DS (m_+1). */
struct DenseQP {
  int n = 0, m = 0;  /* m = m_ = real rows + dummy */
  std::vector<double> Q, D, DU, DS;
};

struct Params {
  wg_herdt_params p;
};

/* --- the QP assembly: generator-vel-ref.cpp + relative-feet-inequalities.cpp + qp-problem.cpp --- */
struct Hull { double X[5], Y[5], A[5], B[5], D[5]; int nv; };

void set_vertices(Hull &h, const wg_herdt_params &P, int foot, int phase, double yaw, bool cop)
{
  /* RelativeFeetInequalities::init_convex_hulls + set_vertices, relative-feet-inequalities.cpp:89-234 */
  if (cop) {
    h.nv = 4;
    const double lxR[4] = {1, 1, -1, -1}, lyR[4] = {-1, 1, 1, -1};
    const double lxL[4] = {1, 1, -1, -1}, lyL[4] = {1, -1, -1, 1};
    const double hw = P.cop_half_x, hh = P.cop_half_y, hhDS = P.cop_half_y + P.ds_feet_distance / 2.0;
    for (int j = 0; j < 4; ++j) {
      if (foot == LEFT) {
        h.X[j] = lxL[j] * hw;
        h.Y[j] = (phase == DS) ? lyL[j] * hhDS - P.ds_feet_distance / 2.0 : lyL[j] * hh;
      } else {
        h.X[j] = lxR[j] * hw;
        h.Y[j] = (phase == DS) ? lyR[j] * hhDS + P.ds_feet_distance / 2.0 : lyR[j] * hh;
      }
    }
  } else {
    h.nv = 5;
    for (int j = 0; j < 5; ++j) {
      h.X[j] = P.foot_hull_x[j];
      h.Y[j] = (foot == LEFT) ? P.foot_hull_y[j] : -P.foot_hull_y[j];
    }
  }
  /* convex_hull_t::rotate(YAW), privatepgtypes.cpp:152-180 */
  for (int j = 0; j < h.nv; ++j) {
    double xo = h.X[j], yo = h.Y[j];
    h.X[j] = (xo * cos(yaw) - yo * sin(yaw));
    h.Y[j] = (xo * sin(yaw) + yo * cos(yaw));
  }
}
void compute_linear_system(Hull &h, int foot)
{
  /* relative-feet-inequalities.cpp:265-319 */
  double sign = (foot == LEFT) ? 1.0 : -1.0;
  for (int i = 0; i < h.nv; ++i) {
    int k = (i + 1 == h.nv) ? 0 : i + 1;
    double y1 = h.Y[i], y2 = h.Y[k], x1 = h.X[i], x2 = h.X[k];
    double dx = y1 - y2, dy = x2 - x1, dc = dx * x1 + dy * y1;
    h.A[i] = sign * dx; h.B[i] = sign * dy; h.D[i] = sign * dc;
  }
}

void build_qp(const wg_herdt_params &P, const wg_herdt_qp_input &in, DenseQP &qp)
{
  Dyn Vel, Cop;
  dyn_velocity(Vel, P.T);
  dyn_cop(Cop, P.T, P.com_height);
  const int ns = in.sup_step[N];
  const int n = 2 * N + 2 * ns;
  const int mreal = 4 * N + 5 * ns;
  const int m = mreal + 1, ld = m + 1;  /* mmax_ = m_ + 1 */
  qp.n = n; qp.m = m;
  qp.Q.assign((size_t)n * n, 0.0); qp.D.assign(n, 0.0);
  qp.DU.assign((size_t)ld * n, 0.0); qp.DS.assign(ld, 0.0);
  auto Q = [&](int i, int j) -> double & { return qp.Q[(size_t)i + (size_t)j * n]; };
  auto DU = [&](int i, int j) -> double & { return qp.DU[(size_t)(i + 1) + (size_t)j * ld]; };  /* row++ */
  auto DSv = [&](int i) -> double & { return qp.DS[i + 1]; };

  /* selection matrices, generator-vel-ref.cpp:138-208 */
  double V[N][WG_HERDT_MAX_STEPS] = {{0}}, VcX[N] = {0}, VcY[N] = {0};
  double Vf[WG_HERDT_MAX_STEPS][WG_HERDT_MAX_STEPS] = {{0}}, VcfX[WG_HERDT_MAX_STEPS] = {0}, VcfY[WG_HERDT_MAX_STEPS] = {0};
  for (int i = 0; i < N; ++i) {
    const int k = i + 1;
    const int sn = in.sup_step[k];
    if (sn > 0) {
      V[i][sn - 1] = 1.0;
      if (sn == 1 && in.sup_changed[k] && in.sup_phase[k] == SS) {
        VcfX[0] = in.sup_x[k - 1]; VcfY[0] = in.sup_y[k - 1];
        Vf[0][0] = 1.0;
      } else if (sn > 1) {
        Vf[sn - 1][sn - 2] = -1.0; Vf[sn - 1][sn - 1] = 1.0;
      }
    } else {
      VcX[i] = in.sup_x[k]; VcY[i] = in.sup_y[k];
    }
  }

  /* build_invariant_part, generator-vel-ref.cpp:588-614 : three terms added one after the other */
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) {
        double s = 0.0, w;
        if (pass == 0) { w = P.w_jerk; s = (i == j) ? 1.0 : 0.0; }          /* U_jerk = I */
        else if (pass == 1) { w = P.w_vel; for (int k = 0; k < N; ++k) s += Vel.U[k][i] * Vel.U[k][j]; }
        else { w = P.w_cop; for (int k = 0; k < N; ++k) s += Cop.U[k][i] * Cop.U[k][j]; }
        s *= w;
        Q(i, j) += s;
        Q(N + i, N + j) += s;
      }
  }
  /* update_problem, generator-vel-ref.cpp:618-674 */
  {
    double Sx[N], Sy[N];
    for (int i = 0; i < N; ++i) {
      double a = 0, b = 0;
      for (int k = 0; k < 3; ++k) { a += Vel.S[i][k] * in.com_x[k]; b += Vel.S[i][k] * in.com_y[k]; }
      Sx[i] = a; Sy[i] = b;
    }
    for (int i = 0; i < N; ++i) {
      double a = 0, b = 0, c = 0, d = 0;
      for (int k = 0; k < N; ++k) {
        a += Vel.U[k][i] * Sx[k]; b += Vel.U[k][i] * Sy[k];
        c += Vel.U[k][i] * in.ref_x[k]; d += Vel.U[k][i] * in.ref_y[k];
      }
      qp.D[i] += a * P.w_vel; qp.D[N + i] += b * P.w_vel;
      qp.D[i] += c * (-P.w_vel); qp.D[N + i] += d * (-P.w_vel);
    }
    /* -a U'V (and transpose), +a V'V */
    for (int i = 0; i < N; ++i)
      for (int s = 0; s < ns; ++s) {
        double t = 0;
        for (int k = 0; k < N; ++k) t += Cop.U[k][i] * V[k][s];
        t *= -P.w_cop;
        Q(i, 2 * N + s) += t; Q(N + i, 2 * N + ns + s) += t;
        Q(2 * N + s, i) += t; Q(2 * N + ns + s, N + i) += t;
      }
    for (int s = 0; s < ns; ++s)
      for (int r = 0; r < ns; ++r) {
        double t = 0;
        for (int k = 0; k < N; ++k) t += V[k][s] * V[k][r];
        t *= P.w_cop;
        Q(2 * N + s, 2 * N + r) += t; Q(2 * N + ns + s, 2 * N + ns + r) += t;
      }
    double Zx[N], Zy[N];
    for (int i = 0; i < N; ++i) {
      double a = 0, b = 0;
      for (int k = 0; k < 3; ++k) { a += Cop.S[i][k] * in.com_x[k]; b += Cop.S[i][k] * in.com_y[k]; }
      Zx[i] = a; Zy[i] = b;
    }
    for (int s = 0; s < ns; ++s) {
      double a = 0, b = 0, c = 0, d = 0;
      for (int k = 0; k < N; ++k) { a += V[k][s] * Zx[k]; b += V[k][s] * Zy[k]; c += V[k][s] * VcX[k]; d += V[k][s] * VcY[k]; }
      qp.D[2 * N + s] += a * (-P.w_cop); qp.D[2 * N + ns + s] += b * (-P.w_cop);
      qp.D[2 * N + s] += c * P.w_cop; qp.D[2 * N + ns + s] += d * P.w_cop;
    }
    /* build_inequalities_cop + build_constraints_cop, generator-vel-ref.cpp:285-316, 394-447 */
    Hull h;
    set_vertices(h, P, in.sup_foot[0], in.sup_phase[0], in.sup_yaw[0], true);
    for (int i = 0; i < N; ++i) {
      const int k = i + 1;
      if (in.sup_changed[k]) set_vertices(h, P, in.sup_foot[k], in.sup_phase[k], in.sup_yaw[k], true);
      compute_linear_system(h, in.sup_foot[k]);
      for (int e = 0; e < 4; ++e) {
        const int row = 4 * i + e;
        for (int j = 0; j < N; ++j) {
          DU(row, j) += -1.0 * (h.A[e] * Cop.U[i][j]);
          DU(row, N + j) += -1.0 * (h.B[e] * Cop.U[i][j]);
        }
        for (int s = 0; s < ns; ++s) {
          DU(row, 2 * N + s) += h.A[e] * V[i][s];
          DU(row, 2 * N + ns + s) += h.B[e] * V[i][s];
        }
        DSv(row) += h.D[e];
        DSv(row) += -1.0 * (h.A[e] * Zx[i]);
        DSv(row) += -1.0 * (h.B[e] * Zy[i]);
        DSv(row) += h.A[e] * VcX[i];
        DSv(row) += h.B[e] * VcY[i];
      }
    }
    /* build_inequalities_feet + build_constraints_feet, generator-vel-ref.cpp:320-354, 450-474 */
    for (int i = 0; i < N; ++i) {
      const int k = i + 1;
      if (in.sup_changed[k] && in.sup_step[k] > 0 && in.sup_phase[k] != DS) {
        Hull f;
        set_vertices(f, P, in.sup_foot[k - 1], in.sup_phase[k - 1], in.sup_yaw[k - 1], false);
        compute_linear_system(f, in.sup_foot[k]);
        const int sn = in.sup_step[k] - 1;
        for (int e = 0; e < 5; ++e) {
          const int row = 4 * N + 5 * sn + e;
          for (int s = 0; s < ns; ++s) {
            DU(row, 2 * N + s) += -1.0 * (f.A[e] * Vf[sn][s]);
            DU(row, 2 * N + ns + s) += -1.0 * (f.B[e] * Vf[sn][s]);
          }
          DSv(row) += f.D[e];
          DSv(row) += f.A[e] * VcfX[sn];
          DSv(row) += f.B[e] * VcfY[sn];
        }
      }
    }
  }
}

int solve_dense(const DenseQP &qp, double *x, double *u /* m + 2n */, int *iters, bool force_textbook)
{
  const int n = qp.n, m = qp.m;
  ql0001_fn ql = force_textbook ? nullptr : load_ql();
  if (ql) {
    int mm = m, me = 0, mmax = m + 1, nn = n, nmax = n, mnn = m + 2 * n, iout = 0, ifail = 0, iprint = 1;
    int lwar = 2 * (3 * n * n / 2 + 10 * n + 2 * (m + 1) + 20000), liwar = 2 * n + 1000;
    std::vector<double> war(lwar), C(qp.Q), d(qp.D), A(qp.DU), b(qp.DS), xl(n, -1e8), xu(n, 1e8);
    std::vector<int> iwar(liwar);
    iwar[0] = 1;
    double eps = 1e-8;
    ql(&mm, &me, &mmax, &nn, &nmax, &mnn, C.data(), d.data(), A.data(), b.data(), xl.data(), xu.data(), x, u, &iout,
       &ifail, &iprint, war.data(), &lwar, iwar.data(), &liwar, &eps);
    if (iters) *iters = -1;
    return ifail;
  }
  /* textbook solver: row-major A without the dummy row */
  std::vector<double> C((size_t)n * n), A((size_t)(m - 1) * n), b(m - 1), uu(m - 1);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) C[(size_t)i * n + j] = qp.Q[(size_t)i + (size_t)j * n];
  for (int i = 0; i < m - 1; ++i) {
    for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = qp.DU[(size_t)(i + 1) + (size_t)j * (m + 1)];
    b[i] = qp.DS[i + 1];
  }
  int rc = oracle_qp_solve(n, m - 1, C.data(), qp.D.data(), A.data(), b.data(), x, uu.data(), iters);
  for (int i = 0; i < m + 2 * n; ++i) u[i] = 0.0;
  for (int i = 0; i < m - 1; ++i) u[i + 1] = uu[i];
  return rc;
}

/* ------------------------------------------------------------------------------------------------
 * Closed loop
 * ---------------------------------------------------------------------------------------------- */
struct Sim {
  wg_herdt_params P;
  /* robot data */
  double lHipL = -30.0 / 180.0 * M_PI, uHipL = 45.0 / 180.0 * M_PI;   /* OrientationsPreview.cpp:52-53 defaults */
  double lHipR = -30.0 / 180.0 * M_PI, uHipR = 45.0 / 180.0 * M_PI;   /* :64-65 (sic: same as left)            */
  double uvLimitFoot = 3.54108;                                          /* |upperVelocityBound| of the hip yaw   */
  double uaLimitHipYaw = 0.1, uLimitFeet = 5.0 / 180.0 * M_PI;         /* :71-73 */
  /* constants of the ctor, ZMPVelocityReferencedQP.cpp:61-118 */
  double TimeBuffer = 0.04, QP_T = 0.1, Ts = 0.005;
  double StepPeriod = 0.8, DSPeriod = 1e9, DSSSPeriod = 0.8;
  unsigned NbStepsSSDS = 2;
  double TSingle = 0.7, TDouble = 0.1, StepHeight = 0.05, FeetDistanceDS = 0.2, Omega = 0.0;
  /* FSM flags, SupportFSM.hh */
  bool InTranslation = false, InRotation = false, PostRotationPhase = false;
  unsigned NbStepsAfterRotation = 0;
  int CurrentSupportFoot = LEFT;
  const double EPS = 1e-6;
  /* state */
  double clock = 0.0, UpperTimeLimitToUpdate = 0.0, TimeToStopOnLineMode = -1.0;
  bool OnLineMode = false, EndingPhase = false, Running = false;
  /* the end-of-walk branch of OnLine (:410-421) was added in 3.1.8 ("Put the CoM at the center of the feet when
   * stopping"); the committed datrefs predate it - tests switch it off to reproduce them */
  bool returnToCentre = true;
  double NewRef[3] = {0, 0, 0}, Ref[3] = {0, 0, 0};
  Support Current;
  double comx[3], comy[3], ComHeight = 0.0;  /* LIPM state m_CoM + m_ComHeight */
  Com TrunkState, TrunkStateT;
  double SupportTimePassed = 0.0, signRotVelTrunk = 1.0;
  std::deque<Zmp> zmpq;
  std::deque<Com> comq;
  std::deque<Foot> lfq, rfq;
  Poly PX{5}, PY{5}, PZ{4}, PTheta{3}, POmega{3}, POmega2{3};
  /* per-QP scratch (Solution_) */
  std::vector<Support> States;
  std::deque<double> SupportAngles, TrunkAngles;
  std::vector<double> Solution;
  int nsol = 0, lastFail = 0, nQP = 0;
  bool textbook = false;
  /* log of every QP handed to the solver */
  bool logging = false;
  std::vector<wg_herdt_qp_input> logIn;
  std::vector<double> logX, logU;  /* 36 / 76 per QP */
  std::vector<int> logMeta;        /* n, m, fail per QP */

  /* SupportFSM::update_vel_reference, SupportFSM.cpp:58-90 */
  void update_vel_reference()
  {
    InTranslation = (fabs(Ref[0]) > 2 * EPS || fabs(Ref[1]) > 2 * EPS);
    if (fabs(Ref[2]) > EPS) {
      InRotation = true;
    } else {
      if (InRotation && !InTranslation) {
        Ref[0] = 2 * EPS; Ref[1] = 2 * EPS;
        if (!PostRotationPhase) {
          CurrentSupportFoot = Current.Foot; NbStepsAfterRotation = 0; PostRotationPhase = true;
        } else {
          if (CurrentSupportFoot != Current.Foot) { CurrentSupportFoot = Current.Foot; ++NbStepsAfterRotation; }
          if (NbStepsAfterRotation > 2) { InRotation = false; PostRotationPhase = false; }
        }
      } else {
        InRotation = false;
      }
    }
  }
  /* SupportFSM::set_support_state, SupportFSM.cpp:94-153 (T_ = QP_T) */
  void set_support_state(double time, unsigned pi, Support &S) const
  {
    const double T = QP_T;
    S.StateChanged = false;
    S.NbInstants++;
    bool given = (fabs(Ref[0]) > EPS || fabs(Ref[1]) > EPS || fabs(Ref[2]) > EPS);
    if (given && S.Phase == DS && (S.TimeLimit - time - EPS) > DSSSPeriod) {
      S.TimeLimit = time + DSSSPeriod - T / 10.0;
      S.NbStepsLeft = NbStepsSSDS;
    }
    if (time + EPS + pi * T >= S.TimeLimit) {
      if (S.Phase == SS && !given && S.NbStepsLeft == 0) {
        S.Phase = DS; S.TimeLimit = time + pi * T + DSPeriod - T / 10.0; S.StateChanged = true; S.NbInstants = 0;
      } else if ((S.Phase == DS && given) || (S.Phase == DS && S.NbStepsLeft > 0)) {
        S.Phase = SS; S.TimeLimit = time + pi * T + StepPeriod - T / 10.0; S.NbStepsLeft = NbStepsSSDS;
        S.StateChanged = true; S.NbInstants = 0;
      } else if ((S.Phase == SS && S.NbStepsLeft > 0) || (S.NbStepsLeft == 0 && given)) {
        S.Foot = (S.Foot == LEFT) ? RIGHT : LEFT;
        S.StateChanged = true; S.NbInstants = 0;
        S.TimeLimit = time + pi * T + StepPeriod - T / 10.0;
        if (pi != 1) ++S.StepNumber;
        if (!given) S.NbStepsLeft = S.NbStepsLeft - 1;
        if (given) S.NbStepsLeft = NbStepsSSDS;
      }
    }
  }
  /* GeneratorVelRef::preview_support_states, generator-vel-ref.cpp:71-134 */
  void preview_support_states(double time)
  {
    States.clear();
    set_support_state(time, 0, Current);
    if (Current.StateChanged) {
      const Foot &F = (Current.Foot == LEFT) ? lfq.front() : rfq.front();
      Current.X = F.x; Current.Y = F.y; Current.Yaw = F.theta * M_PI / 180.0; Current.StartTime = time;
    }
    States.push_back(Current);
    Support Prw = Current;
    Prw.StepNumber = 0;
    for (unsigned pi = 1; pi <= (unsigned)N; ++pi) {
      set_support_state(time, pi, Prw);
      if (Prw.StateChanged) {
        if (pi == 1) {
          const Foot &F = (Prw.Foot == LEFT) ? lfq.back() : rfq.back();
          Prw.X = F.x; Prw.Y = F.y; Prw.Yaw = F.theta * M_PI / 180.0;
          Prw.StartTime = time + pi * Ts;  /* Tprw_ was overwritten by :samplingperiod (mpc-trajectory-generation.cpp:102-106) */
        }
        if (Prw.StepNumber > 0) { Prw.X = 0.0; Prw.Y = 0.0; }
      }
      States.push_back(Prw);
    }
  }

  /* OrientationsPreview, OrientationsPreview.cpp:80-366 */
  bool verify_angle_hip_joint(const Support &CS, double PrwTrunkAngleEnd, double CurrentSupportFootAngle, unsigned StepNumber)
  {
    const double T = QP_T, SSP = StepPeriod;
    double uJ = (CS.Foot == LEFT) ? uHipL : uHipR, lJ = (CS.Foot == LEFT) ? lHipL : lHipR;
    double JointLimit = (TrunkStateT.yaw[1] < 0.0) ? lJ : uJ;
    if (fabs(PrwTrunkAngleEnd - CurrentSupportFootAngle) > fabs(JointLimit)) {
      TrunkStateT.yaw[1] = (CurrentSupportFootAngle + 0.9 * JointLimit - TrunkState.yaw[0] - TrunkState.yaw[1] * T / 2.0) /
                           (SupportTimePassed + StepNumber * SSP - T / 2.0);
      return false;
    }
    return true;
  }
  void preview_orientations(double Time)
  {
    const double T = QP_T, SSP = StepPeriod, EPSo = 0.00000001;
    SupportAngles.clear(); TrunkAngles.clear();
    Support CS = States.front();
    /* verify_acceleration_hip_joint, :254-268 */
    if (CS.Phase != DS) {
      if (fabs(Ref[2] - TrunkState.yaw[1]) > 2.0 / 3.0 * T * uaLimitHipYaw) {
        double sgn = (Ref[2] - TrunkState.yaw[1] < 0.0) ? -1.0 : 1.0;
        TrunkStateT.yaw[1] = TrunkState.yaw[1] + sgn * 2.0 / 3.0 * T * uaLimitHipYaw;
      } else
        TrunkStateT.yaw[1] = Ref[2];
    } else
      TrunkStateT.yaw[1] = 0.0;
    const Foot &LeftFoot = lfq.back(), &RightFoot = rfq.back();
    bool TrunkVelOK = false, TrunkAngleOK = false;
    double FirstFootPreviewed = 0.0;
    signRotVelTrunk = (TrunkStateT.yaw[1] < 0.0) ? -1.0 : 1.0;
    unsigned StepNumber = 0;
    double PreviewedTrunkAngleEnd = 0.0;
    int guard = 0;
    while (!TrunkVelOK) {
      if (++guard > 1000) break;
      double CurrentSupportAngle = (CS.Foot == LEFT) ? lfq[0].theta * M_PI / 180.0 : rfq[0].theta * M_PI / 180.0;
      if (CS.Phase != DS) {
        TrunkAngleOK = false;
        int g2 = 0;
        while (!TrunkAngleOK) {
          if (++g2 > 1000) break;
          if (fabs(TrunkStateT.yaw[1] - TrunkState.yaw[1]) > EPSo) {
            double a = TrunkState.yaw[0], b = TrunkState.yaw[1], c = 0.0;
            double d = 3.0 * (TrunkStateT.yaw[1] - TrunkState.yaw[1]) / (T * T);
            double e = -2.0 * d / (3.0 * T);
            TrunkStateT.yaw[0] = a + b * T + 1.0 / 2.0 * c * T * T + 1.0 / 3.0 * d * T * T * T + 1.0 / 4.0 * e * T * T * T * T;
          } else
            TrunkStateT.yaw[0] = TrunkState.yaw[0] + TrunkState.yaw[1] * T;
          SupportTimePassed = CS.TimeLimit - Time;
          PreviewedTrunkAngleEnd = TrunkStateT.yaw[0] + TrunkStateT.yaw[1] * (SupportTimePassed - T);
          TrunkAngleOK = verify_angle_hip_joint(CS, PreviewedTrunkAngleEnd, CurrentSupportAngle, StepNumber);
        }
      } else {
        SupportTimePassed = CS.TimeLimit + SSP - Time;
        FirstFootPreviewed = 1;
        SupportAngles.push_back(CurrentSupportAngle);
        TrunkStateT.yaw[0] = PreviewedTrunkAngleEnd = TrunkState.yaw[0];
      }
      double PreviousSupportAngle = CurrentSupportAngle;
      double PreviewedSupportFoot = (CS.Foot == LEFT) ? 1.0 : -1.0;
      double CurrentLeftFootAngle = LeftFoot.theta * M_PI / 180.0, CurrentRightFootAngle = RightFoot.theta * M_PI / 180.0;
      for (StepNumber = (unsigned)FirstFootPreviewed; StepNumber <= (unsigned)((int)ceil((N + 1) * T / StepPeriod)); StepNumber++) {
        PreviewedSupportFoot = -PreviewedSupportFoot;
        double PreviewedSupportAngle = PreviewedTrunkAngleEnd + TrunkStateT.yaw[1] * SSP / 2.0;
        /* verify_velocity_hip_joint takes PreviewedSupportAngle BY VALUE (OrientationsPreview.hh): no effect */
        if (PreviewedSupportFoot * (PreviousSupportAngle - PreviewedSupportAngle) - EPSo > uLimitFeet)
          PreviewedSupportAngle = PreviousSupportAngle + signRotVelTrunk * uLimitFeet;
        else if (fabs(PreviewedSupportAngle - PreviousSupportAngle) > uvLimitFoot * SSP)
          PreviewedSupportAngle = PreviousSupportAngle + PreviewedSupportFoot * uvLimitFoot * (SSP - T);
        TrunkAngleOK = verify_angle_hip_joint(CS, PreviewedTrunkAngleEnd, CurrentSupportAngle, StepNumber);
        if (!TrunkAngleOK) { SupportAngles.clear(); TrunkVelOK = false; break; }
        SupportAngles.push_back(PreviewedSupportAngle);
        PreviewedTrunkAngleEnd = PreviewedTrunkAngleEnd + SSP * TrunkStateT.yaw[1];
        PreviousSupportAngle = PreviewedSupportAngle;
        if (PreviewedSupportFoot == 1) CurrentLeftFootAngle = PreviewedSupportAngle;
        else CurrentRightFootAngle = PreviewedSupportAngle;
        TrunkVelOK = true;
      }
      (void)CurrentLeftFootAngle; (void)CurrentRightFootAngle;
    }
    TrunkAngles.push_back(TrunkState.yaw[0]);
    TrunkAngles.push_back(TrunkStateT.yaw[0]);
    for (int i = 1; i < N; ++i) TrunkAngles.push_back(TrunkStateT.yaw[0] + TrunkStateT.yaw[1] * T);
    unsigned j = 0;
    double supportAngle = States[0].Yaw;
    for (int i = 1; i <= N; ++i) {
      if (States[i].StateChanged) { supportAngle = SupportAngles[j]; j++; }
      States[i].Yaw = supportAngle;
    }
  }
  /* OrientationsPreview::interpolate_trunk_orientation, :369-418 */
  void interpolate_trunk_orientation(double Time, int CurrentIndex)
  {
    const double T = QP_T;
    const Support &CS = States.front();
    if (CS.Phase == SS && Time + 3.0 / 2.0 * T < CS.TimeLimit) {
      double a = TrunkState.yaw[1];
      double c = 3.0 * (TrunkStateT.yaw[1] - TrunkState.yaw[1]) / (T * T);
      double d = -2.0 * c / (3.0 * T);
      double Theta = TrunkState.yaw[0];
      comq[CurrentIndex].yaw[0] = TrunkState.yaw[0];
      comq[CurrentIndex].yaw[1] = TrunkState.yaw[1];
      for (int k = 0; k < (int)(T / Ts); k++) {
        double tT = (double)(k + 1) * Ts;
        if (fabs(TrunkStateT.yaw[1] - TrunkState.yaw[1]) - 0.000001 > 0) {
          TrunkState.yaw[0] = (((1.0 / 4.0 * d * tT + 1.0 / 3.0 * c) * tT) * tT + a) * tT + Theta;
          TrunkState.yaw[1] = ((d * tT + c) * tT) * tT + a;
          TrunkState.yaw[2] = (3.0 * d * tT + 2.0 * c) * tT;
        } else
          TrunkState.yaw[0] += Ts * TrunkStateT.yaw[1];
        comq[CurrentIndex + k].yaw[0] = TrunkState.yaw[0];
        comq[CurrentIndex + k].yaw[1] = TrunkState.yaw[1];
      }
    } else if (CS.Phase == DS || Time + 3.0 / 2.0 * T > CS.TimeLimit) {
      for (int k = 0; k < (int)(T / Ts); k++) {
        comq[CurrentIndex + k].yaw[0] = TrunkState.yaw[0];
        comq[CurrentIndex + k].yaw[1] = TrunkState.yaw[1];
      }
    }
  }

  /* LinearizedInvertedPendulum2D::Interpolation, LIPM2D.cpp:157-227 */
  void lipm_interpolation(int CurrentPosition, double CX, double CY)
  {
    int interval = (int)(QP_T / Ts);
    int loopEnd = std::min<int>(interval, ((int)comq.size()) - 1 - CurrentPosition);
    int pos = CurrentPosition;
    const double C2 = -ComHeight / 9.81;
    for (int lk = 0; lk <= loopEnd; lk++, pos++) {
      Com &c = comq[pos];
      double t = (lk + 1) * Ts;
      c.x[0] = comx[0] + t * comx[1] + 0.5 * t * t * comx[2] + t * t * t * CX / 6.0;
      c.x[1] = comx[1] + t * comx[2] + 0.5 * t * t * CX;
      c.x[2] = comx[2] + t * CX;
      c.y[0] = comy[0] + t * comy[1] + 0.5 * t * t * comy[2] + t * t * t * CY / 6.0;
      c.y[1] = comy[1] + t * comy[2] + 0.5 * t * t * CY;
      c.y[2] = comy[2] + t * CY;
      c.yaw[0] = zmpq[pos].theta;
      c.z[0] = ComHeight; c.z[1] = 0; c.z[2] = 0;
      zmpq[pos].px = 1.0 * c.x[0] + 0.0 * c.x[1] + C2 * c.x[2];
      zmpq[pos].py = 1.0 * c.y[0] + 0.0 * c.y[1] + C2 * c.y[2];
    }
  }
  /* LinearizedInvertedPendulum2D::OneIteration, :230-264 (T = QP_T) */
  void lipm_one_iteration(double ux, double uy)
  {
    const double T = QP_T;
    const double A[3][3] = {{1.0, T, T * T / 2.0}, {0.0, 1.0, T}, {0.0, 0.0, 1.0}};
    const double B[3] = {T * T * T / 6.0, T * T / 2.0, T};
    double nx[3], ny[3];
    for (int i = 0; i < 3; ++i) {
      double a = 0, b = 0;
      for (int j = 0; j < 3; ++j) { a += A[i][j] * comx[j]; b += A[i][j] * comy[j]; }
      nx[i] = a + ux * B[i]; ny[i] = b + uy * B[i];
    }
    for (int i = 0; i < 3; ++i) { comx[i] = nx[i]; comy[i] = ny[i]; }
  }

  /* OnLineFootTrajectoryGeneration::UpdateFootPosition, OnLineFootTrajectoryGeneration.cpp:51-199 */
  void update_foot_position(std::deque<Foot> &Sup, std::deque<Foot> &NonSup, int StartIndex, int k,
                            double LocalStart, double UnlockedSwingPeriod, int StepType)
  {
    double InterpolationTime = (double)k * Ts;
    int idx = k + StartIndex;
    double EndOfLiftOff = (TSingle - UnlockedSwingPeriod) * 0.5;
    double StartLanding = EndOfLiftOff + UnlockedSwingPeriod;
    Foot &cur = NonSup[idx];
    const Foot &prev = NonSup[idx - 1];
    Sup[idx] = Sup[StartIndex - 1];
    Sup[idx].stepType = (-1) * StepType;
    cur.stepType = StepType;
    if (LocalStart + InterpolationTime <= EndOfLiftOff || LocalStart + InterpolationTime >= StartLanding) {
      cur.x = prev.x; cur.y = prev.y; cur.theta = prev.theta;
    } else if (LocalStart < EndOfLiftOff && LocalStart + InterpolationTime > EndOfLiftOff) {
      double rt = LocalStart + InterpolationTime - EndOfLiftOff;
      cur.x = PX.val(rt); cur.dx = PX.d1(rt); cur.ddx = PX.d2(rt);
      cur.y = PY.val(rt); cur.dy = PY.d1(rt); cur.ddy = PY.d2(rt);
      cur.theta = PTheta.val(rt); cur.dtheta = PTheta.d1(rt);
    } else {
      cur.x = PX.val(InterpolationTime); cur.dx = PX.d1(InterpolationTime); cur.ddx = PX.d2(InterpolationTime);
      cur.y = PY.val(InterpolationTime); cur.dy = PY.d1(InterpolationTime); cur.ddy = PY.d2(InterpolationTime);
      cur.theta = PTheta.val(InterpolationTime); cur.dtheta = PTheta.d1(InterpolationTime);
    }
    cur.z = PZ.val(LocalStart + InterpolationTime);
    cur.dz = PZ.d1(LocalStart + InterpolationTime);
    if (LocalStart + InterpolationTime < EndOfLiftOff) {
      cur.omega = POmega.val(InterpolationTime);
      cur.domega = POmega.d1(InterpolationTime);
    } else if (LocalStart + InterpolationTime < StartLanding) {
      cur.omega = Omega - POmega2.val(LocalStart + InterpolationTime - EndOfLiftOff) - NonSup[StartIndex - 1].omega2;
    } else {
      cur.omega = POmega.val(LocalStart + InterpolationTime - StartLanding) + NonSup[StartIndex - 1].omega - Omega;
    }
    /* the toe/heel protection terms vanish for omega == 0 (the only case on this path, :omega 0.0) */
  }
  /* interpret_solution + interpolate_feet_positions, :203-346 */
  void interpolate_feet_positions(double Time)
  {
    Support CS = States.front();
    double FPx = 0, FPy = 0;
    if (CS.Phase != DS) {
      unsigned NbStepsPrwd = States.back().StepNumber;
      double Sign = (CS.Foot == LEFT) ? 1.0 : -1.0;
      if (CS.NbStepsLeft > 0 && NbStepsPrwd > 0) {
        FPx = Solution[2 * N]; FPy = Solution[2 * N + NbStepsPrwd];
      } else {
        FPx = CS.X + Sign * sin(CS.Yaw) * FeetDistanceDS;
        FPy = CS.Y - Sign * cos(CS.Yaw) * FeetDistanceDS;
      }
    }
    double LocalInterpolationTime = Time - (CS.TimeLimit - (TDouble + TSingle));
    int StepType = 1;
    unsigned CurrentIndex = lfq.size() - 1;
    lfq.resize((unsigned)(QP_T / Ts) + CurrentIndex + 1);
    rfq.resize((unsigned)(QP_T / Ts) + CurrentIndex + 1);
    if (CS.Phase == SS && Time + 3.0 / 2.0 * QP_T < CS.TimeLimit) {
      double UnlockedSwingPeriod = TSingle * 0.9;
      double EndOfLiftOff = (TSingle - UnlockedSwingPeriod) * 0.5;
      double SwingTimePassed = 0.0;
      if (LocalInterpolationTime > EndOfLiftOff) SwingTimePassed = LocalInterpolationTime - EndOfLiftOff;
      Foot *Last = (CS.Foot == LEFT) ? &rfq[CurrentIndex] : &lfq[CurrentIndex];
      double TimeInterval = UnlockedSwingPeriod - SwingTimePassed;
      poly5_set(PX, TimeInterval, FPx, Last->x, Last->dx, Last->ddx);
      poly5_set(PY, TimeInterval, FPy, Last->y, Last->dy, Last->ddy);
      if (CS.StateChanged) poly4_set(PZ, TSingle, StepHeight);
      poly3_init(PTheta, TimeInterval, SupportAngles[0] * 180.0 / M_PI, Last->theta, Last->dtheta);
      poly3_init(POmega, TimeInterval, 0.0, Last->omega, Last->domega);
      poly3_init(POmega2, TimeInterval, 0.0, Last->omega2, Last->domega2);
      for (int k = 1; k <= (int)(QP_T / Ts); k++) {
        if (CS.Foot == LEFT) update_foot_position(lfq, rfq, CurrentIndex, k, LocalInterpolationTime, UnlockedSwingPeriod, StepType);
        else update_foot_position(rfq, lfq, CurrentIndex, k, LocalInterpolationTime, UnlockedSwingPeriod, StepType);
        lfq[CurrentIndex + k].time = rfq[CurrentIndex + k].time = Time + k * Ts;
      }
    } else if (CS.Phase == DS || Time + 3.0 / 2.0 * QP_T > CS.TimeLimit) {
      for (int k = 0; k <= (int)(QP_T / Ts); k++) {
        rfq[CurrentIndex + k] = rfq[CurrentIndex + k - 1];
        lfq[CurrentIndex + k] = lfq[CurrentIndex + k - 1];
        lfq[CurrentIndex + k].time = rfq[CurrentIndex + k].time = Time + k * Ts;
        lfq[CurrentIndex + k].stepType = rfq[CurrentIndex + k].stepType = 10;
      }
    }
  }

  /* ZMPVelocityReferencedQP::InitOnLine, :213-319 */
  void init(const double com0[3], const double lf0[3], const double rf0[3], const double zmp0[3])
  {
    UpperTimeLimitToUpdate = 0.0;
    OnLineMode = true; EndingPhase = false; TimeToStopOnLineMode = -1.0;
    Foot L, R;
    L.x = lf0[0]; L.y = lf0[1]; L.theta = lf0[2];
    R.x = rf0[0]; R.y = rf0[1]; R.theta = rf0[2];
    int Add = (int)(TimeBuffer / Ts);
    zmpq.assign(Add, Zmp()); comq.assign(Add, Com()); lfq.assign(Add, Foot()); rfq.assign(Add, Foot());
    double t = 0;
    Com start;
    start.x[0] = com0[0]; start.y[0] = com0[1]; start.z[0] = com0[2];
    for (int i = 0; i < Add; ++i) {
      zmpq[i].px = zmp0[0]; zmpq[i].py = zmp0[1]; zmpq[i].pz = zmp0[2]; zmpq[i].theta = 0; zmpq[i].time = t; zmpq[i].stepType = 0;
      comq[i] = start;
      lfq[i] = L; rfq[i] = R;
      lfq[i].time = rfq[i].time = t;
      lfq[i].stepType = rfq[i].stepType = 10;
      t += Ts;
    }
    Current = Support();
    Current.Phase = DS; Current.Foot = LEFT; Current.TimeLimit = 1000000000; Current.NbStepsLeft = 1;
    Current.StateChanged = false; Current.X = L.x; Current.Y = L.y; Current.Yaw = L.theta * M_PI / 180; Current.StartTime = 0.0;
    for (int i = 0; i < 3; ++i) { comx[i] = start.x[i]; comy[i] = start.y[i]; }
    ComHeight = start.z[0];
    TrunkState = start;
    TrunkStateT = Com();
    Running = false;
  }

  void fill_input(wg_herdt_qp_input &in, const double *gx, const double *gy) const
  {
    std::memset(&in, 0, sizeof in);
    for (int i = 0; i < 3; ++i) { in.com_x[i] = comx[i]; in.com_y[i] = comy[i]; }
    for (int i = 0; i < N; ++i) { in.ref_x[i] = gx[i]; in.ref_y[i] = gy[i]; }
    for (int i = 0; i <= N; ++i) {
      in.sup_x[i] = States[i].X; in.sup_y[i] = States[i].Y; in.sup_yaw[i] = States[i].Yaw;
      in.sup_foot[i] = (int8_t)States[i].Foot; in.sup_phase[i] = (int8_t)States[i].Phase;
      in.sup_step[i] = (int8_t)States[i].StepNumber; in.sup_changed[i] = (int8_t)States[i].StateChanged;
    }
  }

  /* ZMPVelocityReferencedQP::OnLine, :324-458 */
  void online(double time)
  {
    if (!OnLineMode) return;
    if (EndingPhase && time >= TimeToStopOnLineMode) OnLineMode = false;
    if (time + 0.00001 > UpperTimeLimitToUpdate) {
      Ref[0] = NewRef[0]; Ref[1] = NewRef[1]; Ref[2] = NewRef[2];
      update_vel_reference();
      preview_support_states(time);
      preview_orientations(time);
      double gx[N], gy[N];
      for (int i = 0; i < N; ++i) {  /* compute_global_reference, generator-vel-ref.cpp:212-229 */
        double yaw = TrunkAngles[i];
        gx[i] = Ref[0] * cos(yaw) - Ref[1] * sin(yaw);
        gy[i] = Ref[1] * cos(yaw) + Ref[0] * sin(yaw);
      }
      wg_herdt_qp_input in;
      fill_input(in, gx, gy);
      DenseQP qp;
      build_qp(P, in, qp);
      std::vector<double> x(qp.n), u(qp.m + 2 * qp.n);
      int iters = 0;
      lastFail = solve_dense(qp, x.data(), u.data(), &iters, textbook);
      Solution = x; nsol = qp.n; ++nQP;
      if (logging) {
        logIn.push_back(in);
        size_t o = logX.size(); logX.resize(o + WG_HERDT_MAX_VARS, 0.0);
        for (int i = 0; i < qp.n; ++i) logX[o + i] = x[i];
        o = logU.size(); logU.resize(o + WG_HERDT_MAX_ROWS + 1, 0.0);
        for (int i = 0; i < qp.m; ++i) logU[o + i] = u[i];
        logMeta.push_back(qp.n); logMeta.push_back(qp.m); logMeta.push_back(lastFail);
      }
      unsigned currentIndex = comq.size();
      comq.resize((unsigned)(QP_T / Ts) + currentIndex);
      zmpq.resize((unsigned)(QP_T / Ts) + currentIndex);
      if (returnToCentre && States.size() && States[0].NbStepsLeft == 0) {
        double jx = (lfq[0].x + rfq[0].x) / 2 - comq[0].x[0];
        double jy = (lfq[0].y + rfq[0].y) / 2 - comq[0].y[0];
        if (fabs(jx) < 1e-3 && fabs(jy) < 1e-3) Running = false;
        const double tf = 0.75;
        jx = 6 / (tf * tf * tf) * (jx - tf * comq[0].x[1] - (tf * tf / 2) * comq[0].x[2]);
        jy = 6 / (tf * tf * tf) * (jy - tf * comq[0].y[1] - (tf * tf / 2) * comq[0].y[2]);
        lipm_interpolation(currentIndex, jx, jy);
        lipm_one_iteration(jx, jy);
      } else {
        Running = true;
        lipm_interpolation(currentIndex, Solution[0], Solution[N]);
        lipm_one_iteration(Solution[0], Solution[N]);
      }
      interpolate_trunk_orientation(time, currentIndex);
      interpolate_feet_positions(time);
      if (!EndingPhase) TimeToStopOnLineMode = UpperTimeLimitToUpdate + QP_T * N;
      UpperTimeLimitToUpdate = UpperTimeLimitToUpdate + QP_T;
    }
  }

  /* one PGI tick: RunOneStepOfTheControlLoop (Herdt branch) + CoMAndFootOnlyStrategy pop.
   * out[37]: datref columns 2..38 (column 1 is the tick time, written by the test harness). */
  int tick(double *out)
  {
    clock += Ts;
    online(clock);
    if (lfq.empty() || rfq.empty() || comq.empty() || zmpq.empty()) return -1;
    Foot L = lfq.front(), R = rfq.front();
    Com c = comq.front();
    Zmp z = zmpq.front();
    lfq.pop_front(); rfq.pop_front(); comq.pop_front(); zmpq.pop_front();
    if (out) {
      int o = 0;
      out[o++] = c.x[0]; out[o++] = c.y[0]; out[o++] = c.z[0]; out[o++] = c.yaw[0];
      out[o++] = c.x[1]; out[o++] = c.y[1]; out[o++] = c.z[1];
      out[o++] = z.px; out[o++] = z.py;
      const Foot *ff[2] = {&L, &R};
      for (int f = 0; f < 2; ++f) {
        const Foot &F = *ff[f];
        out[o++] = F.x; out[o++] = F.y; out[o++] = F.z; out[o++] = F.dx; out[o++] = F.dy; out[o++] = F.dz;
        out[o++] = F.ddx; out[o++] = F.ddy; out[o++] = F.ddz; out[o++] = F.theta; out[o++] = F.omega; out[o++] = F.omega2;
      }
      out[o++] = z.px; out[o++] = z.py; out[o++] = 0.0; out[o++] = 0.0;
    }
    return Running ? 1 : 0;
  }
};

} // namespace

extern "C" {

void oracle_herdt_set_ref_lib(const char *path)
{
  std::snprintf(g_ref_path, sizeof g_ref_path, "%s", path ? path : "");
  g_ql_tried = false; g_ql = nullptr;
}
int oracle_herdt_have_ref_qld(void) { return load_ql() != nullptr; }

/* wg_herdt_default_params restated: FootHalfSize.cpp:62-74 with margins 0.04 (relative-feet-inequalities.cpp:47-49) */
void oracle_herdt_default_params(double sole_length, double sole_width, wg_herdt_params *p)
{
  std::memset(p, 0, sizeof *p);
  p->T = 0.1; p->com_height = 0.814; p->w_jerk = 0.00001; p->w_vel = 1.0; p->w_cop = 0.000001;
  p->cop_half_x = 0.5 * sole_length - 0.04; p->cop_half_y = 0.5 * sole_width - 0.04;
  p->ds_feet_distance = 0.2;
  const double X[5] = {-0.28, -0.2, 0.0, 0.2, 0.28}, Y[5] = {-0.2, -0.3, -0.4, -0.3, -0.2};
  for (int i = 0; i < 5; ++i) { p->foot_hull_x[i] = X[i]; p->foot_hull_y[i] = Y[i]; }
  p->lipm_T = 0.005;
}

/* Dense assembly of one QP in QPProblem's layout.  Q[n*n] col-major, D[n], DU[(m+1)*n] col-major with
 * leading dimension m+1 (row 0 = dummy), DS[m+1].  Buffers must hold the 36/75-sized maxima. */
void oracle_herdt_build_qp(const wg_herdt_params *P, const wg_herdt_qp_input *in, int *n, int *m, double *Q, double *D,
                           double *DU, double *DS)
{
  DenseQP qp;
  build_qp(*P, *in, qp);
  *n = qp.n; *m = qp.m;
  std::memcpy(Q, qp.Q.data(), sizeof(double) * qp.Q.size());
  std::memcpy(D, qp.D.data(), sizeof(double) * qp.D.size());
  std::memcpy(DU, qp.DU.data(), sizeof(double) * qp.DU.size());
  std::memcpy(DS, qp.DS.data(), sizeof(double) * qp.DS.size());
}

/* Build + solve one QP; fills a wg_herdt_qp_output exactly as the product does.
 * solver: 0 = reference QLD if available else textbook, 1 = force textbook. */
int oracle_herdt_solve_qp(const wg_herdt_params *P, const wg_herdt_qp_input *in, wg_herdt_qp_output *out, int solver)
{
  DenseQP qp;
  build_qp(*P, *in, qp);
  std::vector<double> x(qp.n), u(qp.m + 2 * qp.n);
  int iters = 0;
  int fail = solve_dense(qp, x.data(), u.data(), &iters, solver == 1);
  std::memset(out, 0, sizeof *out);
  for (int i = 0; i < qp.n; ++i) out->x[i] = x[i];
  for (int i = 0; i < qp.m; ++i) out->lagr[i] = u[i];
  out->n_vars = qp.n; out->n_rows = qp.m; out->fail = fail; out->iterations = iters;
  /* LIPM OneIteration with T = 0.1 */
  const double T = P->T;
  const double A[3][3] = {{1.0, T, T * T / 2.0}, {0.0, 1.0, T}, {0.0, 0.0, 1.0}};
  const double B[3] = {T * T * T / 6.0, T * T / 2.0, T};
  for (int i = 0; i < 3; ++i) {
    double a = 0, b = 0;
    for (int j = 0; j < 3; ++j) { a += A[i][j] * in->com_x[j]; b += A[i][j] * in->com_y[j]; }
    out->com_next_x[i] = a + x[0] * B[i];
    out->com_next_y[i] = b + x[N] * B[i];
  }
  return fail;
}

long oracle_herdt_solve_qp_batch(const wg_herdt_params *P, int B, const wg_herdt_qp_input *in, wg_herdt_qp_output *out,
                                 int solver)
{
  long nfail = 0;
  for (int b = 0; b < B; ++b) nfail += (oracle_herdt_solve_qp(P, in + b, out + b, solver) != 0);
  return nfail;
}

/* ---- closed-loop simulator (TestHerdt2010 harness) ---- */
void *oracle_herdt_sim_new(const wg_herdt_params *P, const double *com0, const double *lf0, const double *rf0,
                           const double *zmp0, int textbook_solver, int logging)
{
  Sim *s = new Sim();
  s->P = *P;
  s->textbook = textbook_solver != 0;
  s->logging = logging != 0;
  s->init(com0, lf0, rf0, zmp0);
  return s;
}
void oracle_herdt_sim_delete(void *h) { delete static_cast<Sim *>(h); }
void oracle_herdt_sim_set_robot(void *h, double lHipL, double uHipL, double lHipR, double uHipR, double uvLimitFoot)
{
  Sim *s = static_cast<Sim *>(h);
  s->lHipL = lHipL; s->uHipL = uHipL; s->lHipR = lHipR; s->uHipR = uHipR; s->uvLimitFoot = uvLimitFoot;
}
void oracle_herdt_sim_set_return_to_centre(void *h, int on) { static_cast<Sim *>(h)->returnToCentre = on != 0; }
/* Override the initial DS support position.  At the surveyed commit InitOnLine takes it from the left
 * foot (ZMPVelocityReferencedQP.cpp:277-279); the committed datref predates ChangeLog 3.1.8 "Fix PG
 * initialization" and was produced with the support frame at (0, 0.1, 0) - see tests/test_herdt_oracle.py. */
void oracle_herdt_sim_set_initial_support(void *h, double x, double y, double yaw)
{ Sim *s = static_cast<Sim *>(h); s->Current.X = x; s->Current.Y = y; s->Current.Yaw = yaw; }
void oracle_herdt_sim_set_vel_ref(void *h, double x, double y, double yaw)
{ Sim *s = static_cast<Sim *>(h); s->NewRef[0] = x; s->NewRef[1] = y; s->NewRef[2] = yaw; }
/* :numberstepsbeforestop, ZMPVelocityReferencedQP.cpp:197-202 */
void oracle_herdt_sim_steps_before_stop(void *h, unsigned nsteps)
{ Sim *s = static_cast<Sim *>(h); s->Current.NbStepsLeft = nsteps; s->NbStepsSSDS = nsteps; }
void oracle_herdt_sim_stoppg(void *h) { static_cast<Sim *>(h)->EndingPhase = true; }
int oracle_herdt_sim_tick(void *h, double *out37) { return static_cast<Sim *>(h)->tick(out37); }
int oracle_herdt_sim_num_qp(void *h) { return static_cast<Sim *>(h)->nQP; }
int oracle_herdt_sim_last_fail(void *h) { return static_cast<Sim *>(h)->lastFail; }
/* copy the log out: inputs[nqp], x[nqp][36], u[nqp][76], meta[nqp][3] */
int oracle_herdt_sim_get_log(void *h, wg_herdt_qp_input *in, double *x, double *u, int *meta, int capacity)
{
  Sim *s = static_cast<Sim *>(h);
  int n = (int)s->logIn.size();
  if (n > capacity) n = capacity;
  if (in) std::memcpy(in, s->logIn.data(), sizeof(wg_herdt_qp_input) * n);
  if (x) std::memcpy(x, s->logX.data(), sizeof(double) * WG_HERDT_MAX_VARS * n);
  if (u) std::memcpy(u, s->logU.data(), sizeof(double) * (WG_HERDT_MAX_ROWS + 1) * n);
  if (meta) std::memcpy(meta, s->logMeta.data(), sizeof(int) * 3 * n);
  return n;
}

} /* extern "C" */
