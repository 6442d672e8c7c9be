/* oracle/ref_glue_dimitrov.cc - TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" handle onto the reference's own ZMPConstrainedQPFastFormulation object code (Dimitrov2008 generator, PLDP mode),
 * compiled by oracle/Makefile from /root/reference/src/ZMPRefTrajectoryGeneration/ZMPConstrainedQPFastFormulation.cpp,
 * PreviewControl/LinearizedInvertedPendulum2D.cpp and privatepgtypes.cpp where they lie, over the stand-in headers of
 * oracle/ref_shim (dimitrov_prelude.hh replaces the ZMPDiscretization member, which only produces the input buffers, by a
 * do-nothing class; MAL_INVERSE is LAPACK dgetrf_ + dgetri_ as in jrl-mal).  Nothing here restates an algorithm.
 *
 * The reference ends the PROCESS (exit(0), PLDPSolver.cpp:827) when a hot start is infeasible.  oracle/Makefile compiles
 * PLDPSolver.cpp with -Dexit=oracle_ref_exit, so that call lands in oracle_ref_exit() below: while ref_dimitrov_run is on the
 * stack it unwinds to it (the trajectory computed so far is kept and the call reports -100), otherwise it is the real exit().
 */
#include <setjmp.h>
#include <stdlib.h>
#include <unistd.h>
#include <cstring>
#include <deque>
#include <string>
#include <vector>
using std::string;
#include <SimplePluginManager.hh>
#define private public      /* the constants InitConstants() computes are private members; read-only access for the tests */
#define protected public
#include <ZMPRefTrajectoryGeneration/ZMPConstrainedQPFastFormulation.hh>
#undef private
#undef protected

using namespace PatternGeneratorJRL;

namespace {
struct RefDimitrov {
  SimplePluginManager spm;
  CjrlHumanoidDynamicRobot robot;
  ZMPConstrainedQPFastFormulation *gen;
};
struct CwdGuard {     /* the generator and ComputeLinearSystem write debug files into the working directory */
  char old[4096];
  CwdGuard() { if (!getcwd(old, sizeof old)) old[0] = 0; if (chdir("/tmp")) {} }
  ~CwdGuard() { if (old[0] && chdir(old)) {} }
};
void fill(std::deque<FootAbsolutePosition> &q, long n, const double *f, const int *step_type, const double *time)
{
  q.resize(n);
  for (long i = 0; i < n; ++i) {
    std::memset(&q[i], 0, sizeof(FootAbsolutePosition));
    q[i].x = f[4 * i]; q[i].y = f[4 * i + 1]; q[i].z = f[4 * i + 2]; q[i].theta = f[4 * i + 3];
    q[i].stepType = step_type[i]; q[i].time = time[i];
  }
}
jmp_buf g_trap;
volatile int g_trap_armed = 0;
}  // namespace

extern "C" {

void oracle_ref_exit(int code)
{
  if (g_trap_armed) { g_trap_armed = 0; longjmp(g_trap, 1); }
  exit(code);
}

void *ref_dimitrov_new(double sole_length, double sole_width)
{
  CwdGuard g;
  RefDimitrov *h = new RefDimitrov;
  h->robot.left.sole_length = h->robot.right.sole_length = sole_length;
  h->robot.left.sole_width = h->robot.right.sole_width = sole_width;
  h->robot.left.ankle_z = h->robot.right.ankle_z = 0.105;
  h->gen = new ZMPConstrainedQPFastFormulation(&h->spm, "", &h->robot);
  /* the solver's 1.3 ms WALL-CLOCK cap (PLDPSolver.cpp:68-69, :890-900) makes a long solve end at a load-dependent iteration:
   * switched off so that the object is deterministic (the product's cap is an iteration count) */
  if (h->gen->m_PLDPSolver) h->gen->m_PLDPSolver->m_LimitedComputationTime = false;
  return h;
}
void ref_dimitrov_delete(void *hv) { RefDimitrov *h = static_cast<RefDimitrov *>(hv); delete h->gen; delete h; }

/* What InitConstants() left behind (ZMPConstrainedQPFastFormulation.cpp:158-246, 390-680), N = 16:
 * Px [N][3], iPu [N][N], iLQ [2N][2N], OptB [2N][6], OptC [2N][2N] row-major; Pu [N][N] as the reference stores it
 * (m_Pu = iLQ Pu', :660-670); PPu, VPu: the leading [N][N] blocks. */
int ref_dimitrov_constants(void *hv, double *Px, double *iPu, double *iLQ, double *OptB, double *OptC, double *Pu,
                           double *PPu, double *VPu)
{
  RefDimitrov *h = static_cast<RefDimitrov *>(hv);
  ZMPConstrainedQPFastFormulation *g = h->gen;
  const unsigned N = g->m_QP_N;
  for (unsigned i = 0; i < N; ++i) {
    for (unsigned j = 0; j < 3; ++j) Px[3 * i + j] = g->m_Px(i, j);
    for (unsigned j = 0; j < N; ++j) {
      iPu[N * i + j] = g->m_iPu(i, j);
      PPu[N * i + j] = g->m_PPu(i, j);
      VPu[N * i + j] = g->m_VPu(i, j);
    }
  }
  for (unsigned i = 0; i < 2 * N; ++i) {
    for (unsigned j = 0; j < 2 * N; ++j) {
      iLQ[2 * N * i + j] = g->m_iLQ(i, j);
      OptC[2 * N * i + j] = g->m_OptC(i, j);
    }
    for (unsigned j = 0; j < 6; ++j) OptB[6 * i + j] = g->m_OptB(i, j);
  }
  if (g->m_Pu) std::memcpy(Pu, g->m_Pu, sizeof(double) * N * N);
  return (int)N;
}

/* Replace the one constant that comes out of LAPACK (m_iPu = MAL_INVERSE(iLQ Pu'), :678) by a caller-supplied inverse and
 * rebuild the solver object around it exactly as the constructor does (:108-112; PLDPSolver precomputes iPu Px).  The tests use
 * it to show that, given the same inverse, the restated loop is bitwise the reference; with LAPACK's own inverse (1 ulp away)
 * the solver's m_tol pushes (PLDPSolver.cpp:617-618) fall differently and the trajectories agree to 1e-7 only. */
void ref_dimitrov_set_ipu(void *hv, const double *iPu)
{
  RefDimitrov *h = static_cast<RefDimitrov *>(hv);
  ZMPConstrainedQPFastFormulation *g = h->gen;
  const unsigned N = g->m_QP_N;
  for (unsigned i = 0; i < N; ++i)
    for (unsigned j = 0; j < N; ++j) g->m_iPu(i, j) = iPu[N * i + j];
  delete g->m_PLDPSolver;
  g->m_PLDPSolver = new Optimization::Solver::PLDPSolver(N, MAL_RET_MATRIX_DATABLOCK(g->m_iPu), MAL_RET_MATRIX_DATABLOCK(g->m_Px),
                                                         g->m_Pu, MAL_RET_MATRIX_DATABLOCK(g->m_iLQ));
  g->m_PLDPSolver->m_LimitedComputationTime = false;
}

/* BuildZMPTrajectoryFromFootTrajectory (:1097-1520) on caller-supplied buffers.  zmp [n][3] = px, py, theta in/out;
 * com [n][7] = x[0..2], y[0..2], yaw[0] out.  Returns the reference's return value, or -100 when the reference called
 * exit(0) (the buffers then hold what it had computed up to that period). */
int ref_dimitrov_run(void *hv, long n, const double *left, const double *right, const int *step_type, const double *time,
                     double *zmp, double *com, double cx, double cy, double T, unsigned N)
{
  CwdGuard g;
  RefDimitrov *h = static_cast<RefDimitrov *>(hv);
  static std::deque<FootAbsolutePosition> L, R;          /* static: they must survive the longjmp */
  static std::deque<ZMPPosition> Z;
  static std::deque<COMState> Cs;
  fill(L, n, left, step_type, time); fill(R, n, right, step_type, time);
  Z.resize(n); Cs.resize(n);
  for (long i = 0; i < n; ++i) {
    std::memset(&Z[i], 0, sizeof(ZMPPosition));
    std::memset(&Cs[i], 0, sizeof(COMState));
    Z[i].px = zmp[3 * i]; Z[i].py = zmp[3 * i + 1]; Z[i].theta = zmp[3 * i + 2]; Z[i].time = time[i];
  }
  volatile int rc = -100;
  if (setjmp(g_trap) == 0) {
    g_trap_armed = 1;
    rc = h->gen->BuildZMPTrajectoryFromFootTrajectory(L, R, Z, Cs, cx, cy, T, N);
    g_trap_armed = 0;
  }
  for (long i = 0; i < n && i < (long)Z.size(); ++i) {
    zmp[3 * i] = Z[i].px; zmp[3 * i + 1] = Z[i].py;
    double *c = com + 7 * i;
    c[0] = Cs[i].x[0]; c[1] = Cs[i].x[1]; c[2] = Cs[i].x[2]; c[3] = Cs[i].y[0]; c[4] = Cs[i].y[1]; c[5] = Cs[i].y[2];
    c[6] = Cs[i].yaw[0];
  }
  return rc;
}

}  // extern "C"
