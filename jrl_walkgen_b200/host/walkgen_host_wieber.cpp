// walkgen_host_wieber.cpp - class mirror of ZMPQPWithConstraint (Wieber2006) and the ql0001_ entry over the C ABI.
#include "walkgen_host.hh"
#include <sstream>

namespace PatternGeneratorJRL {

static void checkw(int rc, const char *what)
{
  if (rc != WG_OK) throw std::runtime_error(std::string(what) + ": " + wg_last_error(walkgen_b200::default_context()));
}

ZMPQPWithConstraint::ZMPQPWithConstraint(SimplePluginManager *lSPM, std::string, CjrlHumanoidDynamicRobot *aHS)
    : ZMPRefTrajectoryGeneration(lSPM), m_Status(0), m_Done(0)
{
  wg_wieber_default_params(&m_Par);
  if (aHS) {                                   // :240-246: the sole of the robot's feet
    double l = 0.0, w = 0.0;
    aHS->leftFoot()->getSoleSize(l, w);
    m_Par.sole_length = l; m_Par.sole_width = w;
  }
  wg_zmpdisc_default_params(&m_Zd);
  std::string name = ":setpbwconstraint";      // :59-67
  RegisterMethod(name);
}

void ZMPQPWithConstraint::CallMethod(std::string &Method, std::istringstream &strm)   // :1389-1414
{
  if (Method == ":setpbwconstraint") {
    std::string cmd;
    strm >> cmd;
    if (cmd == "XY") strm >> m_Par.constraint_x >> m_Par.constraint_y;
    else if (cmd == "T") strm >> m_Par.T;
    else if (cmd == "N") { unsigned int n = 0; strm >> n; m_Par.N = (int)n; }
  }
  ZMPRefTrajectoryGeneration::CallMethod(Method, strm);
}

void ZMPQPWithConstraint::GetZMPDiscretization(std::deque<ZMPPosition> &ZMPPositions, std::deque<COMState> &COMStates,
                                               std::deque<RelativeFootPosition> &Rel, std::deque<FootAbsolutePosition> &Left,
                                               std::deque<FootAbsolutePosition> &Right, double, COMState &,
                                               MAL_S3_VECTOR_TYPE(double) &, FootAbsolutePosition &InitLeft,
                                               FootAbsolutePosition &InitRight)   // :1340-1387
{
  if (Rel.empty()) return;
  wg_ctx *ctx = walkgen_b200::default_context();
  m_Zd.sampling_period = m_SamplingPeriod; m_Zd.t_single = m_Tsingle; m_Zd.t_double = m_Tdble;
  m_Zd.step_height = m_StepHeight; m_Zd.omega = m_Omega;
  m_Par.sampling_period = m_SamplingPeriod;
  checkw(wg_wieber_set_params(ctx, &m_Par), "wg_wieber_set_params");
  std::vector<wg_rel_step> steps(Rel.size());
  for (size_t i = 0; i < Rel.size(); ++i) {
    std::memset(&steps[i], 0, sizeof(wg_rel_step));
    steps[i].sx = Rel[i].sx; steps[i].sy = Rel[i].sy; steps[i].theta = Rel[i].theta;
    steps[i].ss_time = Rel[i].SStime; steps[i].ds_time = Rel[i].DStime; steps[i].step_type = Rel[i].stepType;
  }
  const int64_t off[2] = {0, (int64_t)steps.size()};
  const double feet[6] = {InitLeft.x, InitLeft.y, InitLeft.theta, InitRight.x, InitRight.y, InitRight.theta};
  wg_kajita_plan *plan = nullptr;
  checkw(wg_kajita_plan_create(ctx, &m_Zd, 1, off, steps.data(), feet, &plan), "wg_kajita_plan_create");
  const int64_t n = wg_kajita_plan_total_samples(plan);
  std::vector<double> com(6 * (size_t)n), zmp(2 * (size_t)n);
  std::vector<wg_foot_sample> l((size_t)n), r((size_t)n);
  int32_t status = 0, done = 0;
  const int rc = wg_wieber_run_batch(ctx, plan, WG_MEM_HOST, com.data(), zmp.data(), l.data(), r.data(), &status, &done, nullptr);
  wg_kajita_plan_destroy(plan);
  checkw(rc, "wg_wieber_run_batch");
  m_Status = status; m_Done = done;
  ZMPPositions.resize((size_t)n); COMStates.resize((size_t)n); Left.resize((size_t)n); Right.resize((size_t)n);
  double t = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    ZMPPosition &z = ZMPPositions[(size_t)i];
    z.px = zmp[2 * i]; z.py = zmp[2 * i + 1]; z.pz = 0.0; z.theta = 0.0; z.time = t; z.stepType = 0;
    COMState &c = COMStates[(size_t)i];
    c.reset();
    for (int k = 0; k < 3; ++k) { c.x[k] = com[6 * i + k]; c.y[k] = com[6 * i + 3 + k]; }
    c.z[0] = m_Par.com_height;
    FootAbsolutePosition *fp[2] = {&Left[(size_t)i], &Right[(size_t)i]};
    const wg_foot_sample *fs[2] = {&l[(size_t)i], &r[(size_t)i]};
    for (int f = 0; f < 2; ++f) {
      std::memset(fp[f], 0, sizeof(FootAbsolutePosition));
      fp[f]->x = fs[f]->x; fp[f]->y = fs[f]->y; fp[f]->z = fs[f]->z; fp[f]->theta = fs[f]->theta;
      fp[f]->omega = fs[f]->omega; fp[f]->omega2 = fs[f]->omega2; fp[f]->time = t;
    }
    t += m_SamplingPeriod;
  }
}

}  // namespace PatternGeneratorJRL

int ql0001_(int *m, int *me, int *mmax, int *n, int *nmax, int *mnn, double *c, double *d, double *a, double *b, double *xl,
            double *xu, double *x, double *u, int *, int *ifail, int *, double *, int *, int *iwar, int *, double *eps1)
{
  if (!m || !me || !mmax || !n || !nmax || !mnn || !c || !d || !x || !ifail) return -1;
  if (iwar && iwar[0] != 1) { *ifail = 5; return 0; }
  wg_ctx *ctx = walkgen_b200::default_context();
  wg_qld_batch q;
  std::memset(&q, 0, sizeof q);
  int32_t mm = *m, mme = *me, fail = 0;
  q.n = *n; q.nmax = *nmax; q.mmax = *mmax; q.shared_hessian = 0;
  q.m = &mm; q.me = &mme; q.C = c; q.d = d; q.A = a; q.a_stride = (long long)*mmax * *n; q.b = b; q.b_stride = *mmax;
  q.xl = xl; q.xu = xu; q.x = x; q.u = u; q.u_stride = *mnn; q.ifail = &fail; q.eps = eps1 ? *eps1 : 0.0;
  if (!xl || !xu) { q.xl = q.xu = nullptr; }
  const int rc = wg_qld_solve_batch(ctx, WG_MEM_HOST, 1, &q);
  *ifail = rc == WG_OK ? fail : 5;
  return 0;
}
