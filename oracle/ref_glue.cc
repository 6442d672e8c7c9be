/* oracle/ref_glue.cc - TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" handles onto the reference's own object code (compiled by oracle/Makefile from
 * /root/reference/src/Mathematics/{qld,OptCholesky,PLDPSolver}.cpp) so that tests and the
 * CPU-baseline leg of bench.py can call it through ctypes.  Nothing here restates an
 * algorithm: every function forwards to the reference class / function named in its comment.
 */
#include <vector>
#include <string>
using std::string;
#include <cstring>
#include <Mathematics/qld.hh>          /* ql0001_            (qld.hh:27-31)        */
#include <Mathematics/OptCholesky.hh>  /* OptCholesky        (OptCholesky.hh:42-140) */
#include <Mathematics/PLDPSolver.hh>   /* Optimization::Solver::PLDPSolver (PLDPSolver.hh:44-68)  */
#include <Mathematics/ConvexHull.hh>   /* ComputeConvexHull::DoComputeConvexHull (ConvexHull.hh)   */
#include <Mathematics/FootConstraintsAsLinearSystem.hh>  /* BuildLinearConstraintInequalities (:60-67) */
#include <deque>

using namespace PatternGeneratorJRL;

extern "C" {

/* ql0001_ has C++ linkage in the reference (qld.hh:27); forward 1:1. */
int ref_ql0001(int *m, int *me, int *mmax, int *n, int *nmax, int *mnn,
               double *c, double *d, double *a, double *b, double *xl, double *xu,
               double *x, double *u, int *iout, int *ifail, int *iprint,
               double *war, int *lwar, int *iwar, int *liwar, double *eps1)
{
  return ql0001_(m, me, mmax, n, nmax, mnn, c, d, a, b, xl, xu, x, u, iout, ifail, iprint,
                 war, lwar, iwar, liwar, eps1);
}

/* Solve `count` QPs of identical (n, m) laid out back to back; used by the CPU baseline so
 * that Python call overhead is outside the timed loop.  Convention is QPProblem::solve's
 * (qp-problem.cpp:246-293): m includes the all-zero dummy row, me=0, mmax=m+1, bounds +-1e8,
 * eps=1e-8, iwar[0]=1.  a is column-major with leading dimension mmax. */
int ref_ql0001_many(int count, int m, int n, const double *C, const double *d,
                    const double *A, const double *b, double *x, double *u, int *ifail_out)
{
  int me = 0, mmax = m + 1, nmax = n, mnn = m + 2 * n, iout = 0, iprint = 1;
  int lwar = 2 * (3 * n * n / 2 + 10 * n + 2 * (m + 1) + 20000), liwar = 2 * n + 1000;
  std::vector<double> war(lwar), xl(n, -1e8), xu(n, 1e8), c(n * n), a(mmax * n), dd(n), bb(mmax);
  std::vector<int> iwar(liwar);
  double eps = 1e-8;
  int nfail = 0;
  for (int k = 0; k < count; ++k) {
    std::memcpy(c.data(), C + (size_t)k * n * n, sizeof(double) * n * n);
    std::memcpy(a.data(), A + (size_t)k * mmax * n, sizeof(double) * mmax * n);
    std::memcpy(dd.data(), d + (size_t)k * n, sizeof(double) * n);
    std::memcpy(bb.data(), b + (size_t)k * mmax, sizeof(double) * mmax);
    int ifail = 0;
    iwar[0] = 1;
    ql0001_(&m, &me, &mmax, &n, &nmax, &mnn, c.data(), dd.data(), a.data(), bb.data(),
            xl.data(), xu.data(), x + (size_t)k * n, u + (size_t)k * mnn, &iout, &ifail, &iprint,
            war.data(), &lwar, iwar.data(), &liwar, &eps);
    if (ifail_out) ifail_out[k] = ifail;
    nfail += (ifail != 0);
  }
  return nfail;
}

/* ---- OptCholesky (OptCholesky.hh:55-102) ---- */
void *ref_optcholesky_new(unsigned nb_max, unsigned card_u, unsigned mode)
{ return new OptCholesky(nb_max, card_u, mode); }
void ref_optcholesky_delete(void *h) { delete static_cast<OptCholesky *>(h); }
void ref_optcholesky_set_A(void *h, double *A, unsigned nb) { static_cast<OptCholesky *>(h)->SetA(A, nb); }
void ref_optcholesky_set_L(void *h, double *L) { static_cast<OptCholesky *>(h)->SetL(L); }
void ref_optcholesky_set_iL(void *h, double *iL) { static_cast<OptCholesky *>(h)->SetiL(iL); }
int ref_optcholesky_add(void *h, unsigned row) { return static_cast<OptCholesky *>(h)->AddActiveConstraint(row); }
int ref_optcholesky_rows(void *h) { return static_cast<OptCholesky *>(h)->CurrentNumberOfRows(); }
void ref_optcholesky_set_to_zero(void *h) { static_cast<OptCholesky *>(h)->SetToZero(); }
int ref_optcholesky_full(void *h) { return static_cast<OptCholesky *>(h)->ComputeNormalCholeskyOnANormal(); }
int ref_optcholesky_inverse(void *h, int mode) { return static_cast<OptCholesky *>(h)->ComputeInverseCholeskyNormal(mode); }

/* ---- PLDPSolver (PLDPSolver.hh:48-68) ---- */
void *ref_pldp_new(unsigned card_u, double *iPu, double *Px, double *Pu, double *iLQ)
{ return new Optimization::Solver::PLDPSolver(card_u, iPu, Px, Pu, iLQ); }
void ref_pldp_delete(void *h) { delete static_cast<Optimization::Solver::PLDPSolver *>(h); }
int ref_pldp_solve(void *h, double *D, unsigned m, double *DPu, double *DPx, double *ZMPRef,
                   double *XkYk, double *X, int *similar, unsigned n_similar,
                   unsigned n_removed, int starting)
{
  std::vector<int> sim(similar, similar + n_similar);
  return static_cast<Optimization::Solver::PLDPSolver *>(h)->SolveProblem(D, m, DPu, DPx, ZMPRef, XkYk, X, sim,
                                                           n_removed, starting != 0);
}

/* ---- ComputeConvexHull::DoComputeConvexHull (ConvexHull.cpp:87-203): points as (col=x, row=y) pairs ---- */
int ref_convex_hull(int n, const double *xy, double *hull_xy, int cap)
{
  std::vector<CH_Point> in(n), out;
  for (int i = 0; i < n; ++i) { in[i].col = xy[2 * i]; in[i].row = xy[2 * i + 1]; }
  ComputeConvexHull ch;
  ch.DoComputeConvexHull(in, out);
  for (int i = 0; i < (int)out.size() && i < cap; ++i) { hull_xy[2 * i] = out[i].col; hull_xy[2 * i + 1] = out[i].row; }
  return (int)out.size();
}

/* ---- FootConstraintsAsLinearSystem::BuildLinearConstraintInequalities (FootConstraintsAsLinearSystem.cpp:258-539)
 * feet: [n][4] = x, y, z, theta(deg); step_type/time: the left foot's, [n].  Output polygon p: rows[p], A[p][8][2],
 * Bv[p][8], center[p][2], similar[p][8], t0t1[p][2].  Returns the number of polygons (or -1). */
int ref_fcals_build(int n, const double *left, const double *right, const int *step_type, const double *time,
                    double sole_length, double sole_width, double cx, double cy, int cap, int *rows, double *A,
                    double *Bv, double *center, int *similar, double *t0t1)
{
  CjrlHumanoidDynamicRobot robot;
  robot.left.sole_length = robot.right.sole_length = sole_length;
  robot.left.sole_width = robot.right.sole_width = sole_width;
  robot.left.ankle_z = robot.right.ankle_z = 0.105;
  SimplePluginManager spm;
  FootConstraintsAsLinearSystem fcals(&spm, &robot);
  std::deque<FootAbsolutePosition> L(n), R(n);
  for (int i = 0; i < n; ++i) {
    std::memset(&L[i], 0, sizeof(FootAbsolutePosition)); std::memset(&R[i], 0, sizeof(FootAbsolutePosition));
    L[i].x = left[4 * i]; L[i].y = left[4 * i + 1]; L[i].z = left[4 * i + 2]; L[i].theta = left[4 * i + 3];
    R[i].x = right[4 * i]; R[i].y = right[4 * i + 1]; R[i].z = right[4 * i + 2]; R[i].theta = right[4 * i + 3];
    L[i].stepType = step_type[i]; L[i].time = time[i]; R[i].time = time[i];
  }
  std::deque<LinearConstraintInequality_t *> q;
  int rc = fcals.BuildLinearConstraintInequalities(L, R, q, cx, cy);
  int np = (int)q.size();
  for (int p = 0; p < np; ++p) {
    if (p < cap) {
      const int nr = (int)q[p]->A.size1();
      rows[p] = nr;
      for (int j = 0; j < nr && j < 8; ++j) {
        A[(p * 8 + j) * 2] = q[p]->A(j, 0); A[(p * 8 + j) * 2 + 1] = q[p]->A(j, 1);
        Bv[p * 8 + j] = q[p]->B(j, 0);
        similar[p * 8 + j] = q[p]->SimilarConstraints[j];
      }
      center[2 * p] = q[p]->Center(0); center[2 * p + 1] = q[p]->Center(1);
      t0t1[2 * p] = q[p]->StartingTime; t0t1[2 * p + 1] = q[p]->EndingTime;
    }
    delete q[p];
  }
  return rc < 0 ? rc : np;
}

} /* extern "C" */
