/* oracle/ref_glue_wieber.cc - TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" handle onto the reference's own ZMPQPWithConstraint object code (Wieber2006 generator), compiled by
 * oracle/Makefile from /root/reference/src/ZMPRefTrajectoryGeneration/ZMPQPWithConstraint.cpp where it lies, over the
 * stand-in headers of oracle/ref_shim (wieber_prelude.hh replaces the ZMPDiscretization member, which only produces the
 * input buffers, by a do-nothing class).  Nothing here restates an algorithm.
 */
#include <unistd.h>
#include <cstring>
#include <deque>
#include <string>
#include <vector>
using std::string;
#include <SimplePluginManager.hh>
#include <ZMPRefTrajectoryGeneration/ZMPQPWithConstraint.hh>

using namespace PatternGeneratorJRL;

namespace {
struct RefWieber {
  SimplePluginManager spm;
  CjrlHumanoidDynamicRobot robot;
  ZMPQPWithConstraint *gen;
};
struct CwdGuard {     /* ComputeLinearSystem appends to "Constraints.dat" in the working directory (:107-121) */
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
}  // namespace

extern "C" {

void *ref_wieber_new(double sole_length, double sole_width)
{
  RefWieber *h = new RefWieber;
  h->robot.left.sole_length = h->robot.right.sole_length = sole_length;
  h->robot.left.sole_width = h->robot.right.sole_width = sole_width;
  h->robot.left.ankle_z = h->robot.right.ankle_z = 0.105;
  h->gen = new ZMPQPWithConstraint(&h->spm, "", &h->robot);
  return h;
}
void ref_wieber_delete(void *hv) { RefWieber *h = static_cast<RefWieber *>(hv); delete h->gen; delete h; }

/* BuildLinearConstraintInequalities (:229-502): polygons out as rows [np][cap_rows][3] = A0, A1, B; times [np][2]. */
int ref_wieber_polygons(void *hv, long n, const double *left, const double *right, const int *step_type, const double *time,
                        double cx, double cy, int cap, int cap_rows, double *rows, double *times, int *nrows)
{
  CwdGuard g;
  RefWieber *h = static_cast<RefWieber *>(hv);
  std::deque<FootAbsolutePosition> L, R;
  fill(L, n, left, step_type, time); fill(R, n, right, step_type, time);
  std::deque<LinearConstraintInequality_t *> Q;
  h->gen->BuildLinearConstraintInequalities(L, R, Q, cx, cy);
  const int np = (int)Q.size();
  for (int p = 0; p < np && p < cap; ++p) {
    nrows[p] = (int)MAL_MATRIX_NB_ROWS(Q[p]->A);
    for (int j = 0; j < nrows[p] && j < cap_rows; ++j) {
      rows[((size_t)p * cap_rows + j) * 3] = Q[p]->A(j, 0);
      rows[((size_t)p * cap_rows + j) * 3 + 1] = Q[p]->A(j, 1);
      rows[((size_t)p * cap_rows + j) * 3 + 2] = Q[p]->B(j, 0);
    }
    times[2 * p] = Q[p]->StartingTime; times[2 * p + 1] = Q[p]->EndingTime;
  }
  for (int p = 0; p < np; ++p) delete Q[p];
  return np;
}

/* BuildZMPTrajectoryFromFootTrajectory (:665-1338) on caller-supplied buffers.  zmp [n][3] = px, py, theta in/out;
 * com [n][7] = x[0..2], y[0..2], yaw[0] out.  Returns the reference's return value (0, or -1 on IFAIL / violated row). */
int ref_wieber_run(void *hv, long n, const double *left, const double *right, const int *step_type, const double *time,
                   double *zmp, double *com, double cx, double cy, double T, unsigned N)
{
  CwdGuard g;
  RefWieber *h = static_cast<RefWieber *>(hv);
  std::deque<FootAbsolutePosition> L, R;
  fill(L, n, left, step_type, time); fill(R, n, right, step_type, time);
  std::deque<ZMPPosition> Z(n);
  std::deque<COMState> Cs(n);
  for (long i = 0; i < n; ++i) {
    std::memset(&Z[i], 0, sizeof(ZMPPosition));
    std::memset(&Cs[i], 0, sizeof(COMState));
    Z[i].px = zmp[3 * i]; Z[i].py = zmp[3 * i + 1]; Z[i].theta = zmp[3 * i + 2]; Z[i].time = time[i];
  }
  const int rc = h->gen->BuildZMPTrajectoryFromFootTrajectory(L, R, Z, Cs, cx, cy, T, N);
  for (long i = 0; i < n; ++i) {
    zmp[3 * i] = Z[i].px; zmp[3 * i + 1] = Z[i].py;
    double *c = com + 7 * i;
    c[0] = Cs[i].x[0]; c[1] = Cs[i].x[1]; c[2] = Cs[i].x[2]; c[3] = Cs[i].y[0]; c[4] = Cs[i].y[1]; c[5] = Cs[i].y[2];
    c[6] = Cs[i].yaw[0];
  }
  return rc;
}

}  // extern "C"
