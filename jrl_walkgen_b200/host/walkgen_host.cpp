// walkgen_host.cpp - implementation of the host-side class mirror (walkgen_host.hh) over the C ABI.
// No algorithm lives here: every numerical result comes from libwalkgen_b200's CUDA kernels.
#include "walkgen_host.hh"
#include <fstream>
#include <iostream>
#include <cmath>
#include <cstdlib>
#include <mutex>

namespace walkgen_b200 {

wg_ctx *default_context()
{
  static wg_ctx *ctx = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!ctx) {
    const char *dev = std::getenv("WG_DEVICE");
    const int rc = wg_ctx_create(dev ? std::atoi(dev) : 0, &ctx);
    if (rc != WG_OK || !ctx) {
      ctx = nullptr;
      throw std::runtime_error("walkgen_b200: no usable CUDA device (there is no CPU fallback)");
    }
  }
  return ctx;
}

void walkgen_b200_check(int rc, const char *what)
{
  if (rc != WG_OK) throw std::runtime_error(std::string("walkgen_b200: ") + what + ": " + wg_last_error(walkgen_b200::default_context()));
}

static void check(int rc, const char *what)
{
  if (rc != WG_OK) throw std::runtime_error(std::string("walkgen_b200: ") + what + ": " + wg_last_error(default_context()));
}

}  // namespace walkgen_b200

using walkgen_b200::check;
using walkgen_b200::default_context;

namespace PatternGeneratorJRL {

// ---------------------------------------------------------------------------------------------
// SimplePlugin / SimplePluginManager
// ---------------------------------------------------------------------------------------------
SimplePlugin::~SimplePlugin()
{
  if (m_SimplePluginManager) m_SimplePluginManager->UnregisterPlugin(this);
}
bool SimplePlugin::RegisterMethod(std::string &MethodName)
{
  return m_SimplePluginManager ? m_SimplePluginManager->RegisterMethod(MethodName, this) : false;
}
bool SimplePluginManager::RegisterMethod(std::string &MethodName, SimplePlugin *aSP)
{
  m_SimplePlugins.insert(std::pair<std::string, SimplePlugin *>(MethodName, aSP));
  return true;
}
void SimplePluginManager::UnregisterPlugin(SimplePlugin *aSP)
{
  for (auto it = m_SimplePlugins.begin(); it != m_SimplePlugins.end();) {
    if (it->second == aSP) it = m_SimplePlugins.erase(it);
    else ++it;
  }
}
bool SimplePluginManager::CallMethod(std::string &MethodName, std::istringstream &istrm)
{
  // the rest of the buffer is handed, from its start, to every plugin registered under the name
  std::string rest;
  std::getline(istrm, rest, '\0');
  bool found = false;
  auto range = m_SimplePlugins.equal_range(MethodName);
  std::vector<SimplePlugin *> targets;
  for (auto it = range.first; it != range.second; ++it) targets.push_back(it->second);
  for (SimplePlugin *sp : targets) {
    if (!sp) continue;
    std::istringstream iss(rest);
    sp->CallMethod(MethodName, iss);
    found = true;
  }
  return found;
}

// ---------------------------------------------------------------------------------------------
// PreviewControl
// ---------------------------------------------------------------------------------------------
static const PreviewControl *g_gains_owner = nullptr;   // whose gains are loaded in the context

PreviewControl::PreviewControl(SimplePluginManager *lSPM, unsigned int defaultMode, bool computeWeightsAutomatically)
    : SimplePlugin(lSPM), m_SamplingPeriod(0.0), m_PreviewControlTime(0.0), m_Zc(0.0), m_Coherent(false),
      m_AutoComputeWeights(computeWeightsAutomatically), m_DefaultWeightComputationMode(defaultMode),
      m_SizeOfPreviewWindow(0)
{
  std::memset(&m_Gains, 0, sizeof m_Gains);
  std::string names[3] = {":samplingperiod", ":previewcontroltime", ":comheight"};
  for (auto &n : names) RegisterMethod(n);
}
PreviewControl::~PreviewControl()
{
  if (g_gains_owner == this) g_gains_owner = nullptr;
}
void PreviewControl::SetSamplingPeriod(double v)
{
  if (m_SamplingPeriod != v) m_Coherent = false;
  m_SamplingPeriod = v;
  if (m_AutoComputeWeights) ComputeOptimalWeights(m_DefaultWeightComputationMode);
}
void PreviewControl::SetPreviewControlTime(double v)
{
  if (m_PreviewControlTime != v) m_Coherent = false;
  m_PreviewControlTime = v;
  if (m_AutoComputeWeights) ComputeOptimalWeights(m_DefaultWeightComputationMode);
}
void PreviewControl::SetHeightOfCoM(double v)
{
  if (m_Zc != v) m_Coherent = false;
  m_Zc = v;
  if (m_AutoComputeWeights) ComputeOptimalWeights(m_DefaultWeightComputationMode);
}
void PreviewControl::ComputeOptimalWeights(unsigned int mode)
{
  // the reference solves the Riccati equation whatever the parameters; incomplete parameter sets (the plugin
  // commands arrive one by one) simply leave the object incoherent here
  if (!(m_SamplingPeriod > 0.0) || !(m_PreviewControlTime > 0.0) || !(m_Zc > 0.0)) return;
  if (wg_preview_gains(m_SamplingPeriod, m_PreviewControlTime, m_Zc, (int)mode, &m_Gains) != WG_OK) return;
  m_SizeOfPreviewWindow = (unsigned)m_Gains.NL;
  m_Coherent = true;
  if (g_gains_owner == this) g_gains_owner = nullptr;   // force a reload
}
void PreviewControl::ReadPrecomputedFile(std::string aFileName)
{
  std::ifstream aif(aFileName.c_str(), std::ifstream::in);
  if (!aif.is_open()) {
    std::cerr << "PreviewControl - Unable to open " << aFileName << std::endl;
    return;
  }
  wg_preview_gains_t g;
  std::memset(&g, 0, sizeof g);
  aif >> g.zc >> g.T >> g.preview_time;
  float r;                                    // the reference parses the gains into a float (:157)
  for (int i = 0; i < 3; ++i) { aif >> r; g.Kx[i] = r; }
  aif >> r; g.Ks = r;
  const unsigned NL = (unsigned)(g.preview_time / g.T);
  if (NL == 0 || NL > WG_PREVIEW_MAX_NL) {
    std::cerr << "PreviewControl - window of " << NL << " samples not supported" << std::endl;
    return;
  }
  for (unsigned i = 0; i < NL; ++i) { aif >> r; g.F[i] = r; }
  const double T = g.T;
  const double A[9] = {1.0, T, T * T / 2.0, 0.0, 1.0, T, 0.0, 0.0, 1.0};
  std::memcpy(g.A, A, sizeof A);
  g.B[0] = T * T * T / 6.0; g.B[1] = T * T / 2.0; g.B[2] = T;
  g.C[0] = 1.0; g.C[1] = 0.0; g.C[2] = -g.zc / 9.81;
  g.NL = (int)NL;
  g.mode = (int)m_DefaultWeightComputationMode;
  m_Gains = g;
  m_Zc = g.zc; m_SamplingPeriod = g.T; m_PreviewControlTime = g.preview_time;
  m_SizeOfPreviewWindow = NL;
  m_Coherent = true;
  if (g_gains_owner == this) g_gains_owner = nullptr;   // force a reload
}
void PreviewControl::CallMethod(std::string &Method, std::istringstream &strm)
{
  double v;
  if (Method == ":samplingperiod") { if (strm.good()) { strm >> v; SetSamplingPeriod(v); } }
  else if (Method == ":previewcontroltime") { if (strm.good()) { strm >> v; SetPreviewControlTime(v); } }
  else if (Method == ":comheight") { if (strm.good()) { strm >> v; SetHeightOfCoM(v); } }
  else if (Method == ":computeweightsofpreview") {
    std::string initialpos;
    if (strm.good()) {
      strm >> initialpos;
      if (initialpos == "withinitialpos") ComputeOptimalWeights(OptimalControllerSolver::MODE_WITH_INITIALPOS);
      else if (initialpos == "withoutinitialpos") ComputeOptimalWeights(OptimalControllerSolver::MODE_WITHOUT_INITIALPOS);
    }
  }
}
static void load_gains(const PreviewControl *pc, const wg_preview_gains_t &g)
{
  if (g_gains_owner != pc) {
    check(wg_preview_set_gains(default_context(), &g), "wg_preview_set_gains");
    g_gains_owner = pc;
  }
}
int PreviewControl::OneIterationOfPreview(MAL_MATRIX(&x, double), MAL_MATRIX(&y, double), double &sxzmp, double &syzmp,
                                          std::deque<ZMPPosition> &ZMPPositions, unsigned int lindex, double &zmpx2,
                                          double &zmpy2, bool Simulation)
{
  if (!m_Coherent) throw std::runtime_error("PreviewControl: weights not computed");
  if (ZMPPositions.size() < m_SizeOfPreviewWindow || ZMPPositions.size() - lindex < m_SizeOfPreviewWindow)
    throw std::runtime_error("ZMPPositions.size()<m_SizeOfPreviewWindow:");   // LTHROW, PreviewControl.cpp:341-344
  load_gains(this, m_Gains);
  std::vector<double> w(2 * (size_t)m_SizeOfPreviewWindow);
  for (unsigned i = 0; i < m_SizeOfPreviewWindow; ++i) {
    w[2 * i] = ZMPPositions[lindex + i].px;
    w[2 * i + 1] = ZMPPositions[lindex + i].py;
  }
  double xs[3] = {x(0, 0), x(1, 0), x(2, 0)}, ys[3] = {y(0, 0), y(1, 0), y(2, 0)};
  check(wg_preview_one_iteration(default_context(), xs, ys, &sxzmp, &syzmp, w.data(), (int)m_SizeOfPreviewWindow, &zmpx2,
                                 &zmpy2, Simulation ? 1 : 0), "wg_preview_one_iteration");
  for (int i = 0; i < 3; ++i) { x(i, 0) = xs[i]; y(i, 0) = ys[i]; }
  return 0;
}
int PreviewControl::run1d(walkgen_b200::Matrix &x, double &sxzmp, const std::vector<double> &window, double &zmpx2,
                          bool Simulation)
{
  load_gains(this, m_Gains);
  std::vector<double> w(2 * window.size(), 0.0);
  for (size_t i = 0; i < window.size(); ++i) w[2 * i] = window[i];
  double xs[3] = {x(0, 0), x(1, 0), x(2, 0)}, ys[3] = {0, 0, 0}, sy = 0.0, zy = 0.0;
  check(wg_preview_one_iteration(default_context(), xs, ys, &sxzmp, &sy, w.data(), (int)window.size(), &zmpx2, &zy,
                                 Simulation ? 1 : 0), "wg_preview_one_iteration");
  for (int i = 0; i < 3; ++i) x(i, 0) = xs[i];
  return 0;
}
int PreviewControl::OneIterationOfPreview1D(MAL_MATRIX(&x, double), double &sxzmp, std::deque<double> &ZMPPositions,
                                            unsigned int lindex, double &zmpx2, bool Simulation)
{
  if (!m_Coherent) throw std::runtime_error("PreviewControl: weights not computed");
  // the reference exit(0)s here (PreviewControl.cpp:394-399); an exception is the library-safe equivalent
  if (ZMPPositions.size() < m_SizeOfPreviewWindow || ZMPPositions.size() - lindex < m_SizeOfPreviewWindow)
    throw std::runtime_error("ZMPPositions.size()< m_SizeOfPreviewWindow");
  std::vector<double> w(ZMPPositions.begin() + lindex, ZMPPositions.begin() + lindex + m_SizeOfPreviewWindow);
  return run1d(x, sxzmp, w, zmpx2, Simulation);
}
int PreviewControl::OneIterationOfPreview1D(MAL_MATRIX(&x, double), double &sxzmp, std::vector<double> &Z,
                                            unsigned int lindex, double &zmpx2, bool Simulation)
{
  if (!m_Coherent) throw std::runtime_error("PreviewControl: weights not computed");
  if (Z.size() < m_SizeOfPreviewWindow) throw std::runtime_error("ZMPPositions.size()< m_SizeOfPreviewWindow");
  const unsigned NL = m_SizeOfPreviewWindow;
  const int TestSize = (int)Z.size() - (int)lindex - (int)NL;
  std::vector<double> w(NL, 0.0);
  if (TestSize >= 0) {
    for (unsigned i = 0; i < NL; ++i) w[i] = Z[lindex + i];
    return run1d(x, sxzmp, w, zmpx2, Simulation);
  }
  // wrap-around branch, PreviewControl.cpp:455-466: the reference indexes F with the ABSOLUTE buffer index
  // (ux += F(i) Z[i] for i in [lindex, size) and again for i in [0, StillToRealized)); reproduced as written
  for (unsigned i = lindex; i < Z.size() && i < NL; ++i) w[i] += Z[i];
  const int Still = (int)NL - (int)Z.size() + (int)lindex;
  for (int i = 0; i < Still && i < (int)NL; ++i) w[i] += Z[i];
  int rc = run1d(x, sxzmp, w, zmpx2, false);
  if (Simulation) sxzmp += (Z[lindex] - zmpx2);
  return rc;
}
void PreviewControl::BindGains() const
{
  if (!m_Coherent) throw std::runtime_error("PreviewControl: weights not computed");
  load_gains(this, m_Gains);
}
int PreviewControl::RunWholeTrajectory(const std::deque<ZMPPosition> &Z, MAL_MATRIX(&x, double), MAL_MATRIX(&y, double),
                                       double &sxzmp, double &syzmp, std::vector<double> &com6, std::vector<double> &zmp2,
                                       bool Simulation)
{
  if (!m_Coherent) throw std::runtime_error("PreviewControl: weights not computed");
  load_gains(this, m_Gains);
  wg_ctx *ctx = default_context();
  const int64_t offs[2] = {0, (int64_t)Z.size()};
  std::vector<double> w(2 * Z.size());
  for (size_t i = 0; i < Z.size(); ++i) { w[2 * i] = Z[i].px; w[2 * i + 1] = Z[i].py; }
  wg_preview_plan *plan = nullptr;
  check(wg_preview_plan_create(ctx, 1, offs, &plan), "wg_preview_plan_create");
  double st[8] = {x(0, 0), x(1, 0), x(2, 0), y(0, 0), y(1, 0), y(2, 0), sxzmp, syzmp};
  com6.assign(6 * Z.size(), 0.0); zmp2.assign(2 * Z.size(), 0.0);
  const int rc = wg_preview_run_batch(ctx, plan, WG_MEM_HOST, w.data(), st, com6.data(), zmp2.data(), Simulation ? 1 : 0);
  const int64_t steps = wg_preview_plan_total_steps(plan);
  wg_preview_plan_destroy(plan);
  check(rc, "wg_preview_run_batch");
  for (int i = 0; i < 3; ++i) { x(i, 0) = st[i]; y(i, 0) = st[3 + i]; }
  sxzmp = st[6]; syzmp = st[7];
  return (int)steps;
}

// ---------------------------------------------------------------------------------------------
// OptCholesky
// ---------------------------------------------------------------------------------------------
OptCholesky::OptCholesky(unsigned int lNbMaxOfConstraints, unsigned int lCardU, unsigned int mode)
    : m_NbMaxOfConstraints(lNbMaxOfConstraints), m_CardU(lCardU), m_A(0), m_L(0), m_iL(0), m_UpdateMode(mode),
      m_NbOfConstraints(0) {}
OptCholesky::~OptCholesky() {}
void OptCholesky::SetToZero() { m_SetActiveConstraints.clear(); }
void OptCholesky::SetA(double *aA, unsigned int n) { m_A = aA; m_NbOfConstraints = n; }
void OptCholesky::SetL(double *aL) { m_L = aL; }
void OptCholesky::SetiL(double *aiL) { m_iL = aiL; }
int OptCholesky::CurrentNumberOfRows() { return (int)m_SetActiveConstraints.size(); }
int OptCholesky::AddActiveConstraints(std::vector<unsigned int> &l)
{
  int r = 0;
  for (unsigned li = 0; li < l.size(); ++li) {
    r = AddActiveConstraint(l[li]);
    if (r < 0) return -((int)li);
  }
  return r;
}
int OptCholesky::AddActiveConstraint(unsigned int aConstraint)
{
  m_SetActiveConstraints.push_back(aConstraint);
  if (m_A == 0 || m_L == 0) return 0;   // UpdateCholeskyMatrix* returns -1 but AddActiveConstraint ignores it (:92-104)
  const int k = (int)m_SetActiveConstraints.size();
  std::vector<int32_t> rows(m_SetActiveConstraints.begin(), m_SetActiveConstraints.end());
  const long long a_elems = (m_UpdateMode == MODE_FORTRAN) ? (long long)(m_NbOfConstraints + 1) * m_CardU
                                                            : (long long)m_NbOfConstraints * m_CardU;
  const int rc = wg_optcholesky_add_rows_batch(default_context(), WG_MEM_HOST, 1, (int)m_UpdateMode,
                                               (int)m_NbMaxOfConstraints, (int)m_CardU, (int)m_NbOfConstraints, m_A,
                                               a_elems, rows.data(), k, k - 1, k, m_L,
                                               (long long)m_NbMaxOfConstraints * m_NbMaxOfConstraints);
  check(rc, "wg_optcholesky_add_rows_batch");
  return 0;
}
int OptCholesky::ComputeNormalCholeskyOnANormal()
{
  if (m_A == 0 || m_L == 0) return -1;
  if (m_NbMaxOfConstraints != m_CardU) return -2;
  check(wg_optcholesky_full_batch(default_context(), WG_MEM_HOST, 1, (int)m_NbMaxOfConstraints, m_A, m_L, nullptr, 0),
        "wg_optcholesky_full_batch");
  return 0;
}
int OptCholesky::ComputeInverseCholeskyNormal(int mode)
{
  if (m_iL == 0) return -1;
  const int size = (mode == 0) ? (int)m_SetActiveConstraints.size() : (int)m_NbMaxOfConstraints;
  check(wg_optcholesky_full_batch(default_context(), WG_MEM_HOST, 1, (int)m_NbMaxOfConstraints, nullptr, m_L, m_iL, size),
        "wg_optcholesky_full_batch");
  return 0;
}

}  // namespace PatternGeneratorJRL

// ---------------------------------------------------------------------------------------------
// PLDPSolver
// ---------------------------------------------------------------------------------------------
namespace Optimization {
namespace Solver {

PLDPSolver::PLDPSolver(unsigned int CardU, double *iPu, double *Px, double *Pu, double *)
    : m_CardV(CardU)
{
  std::memset(&m_Hot, 0, sizeof m_Hot);
  std::memset(&m_Info, 0, sizeof m_Info);
  check(wg_pldp_set_constants(default_context(), (int)CardU, iPu, Px, Pu), "wg_pldp_set_constants");
}
PLDPSolver::~PLDPSolver() {}
int PLDPSolver::SolveProblem(double *D, unsigned int NbOfConstraints, double *DPu, double *DPx, double *ZMPRef,
                             double *XkYk, double *X, std::vector<int> &SimilarConstraint,
                             unsigned int NumberOfRemovedConstraints, bool StartingSequence)
{
  wg_pldp_batch b;
  std::memset(&b, 0, sizeof b);
  int32_t m = (int32_t)NbOfConstraints, nrem = (int32_t)NumberOfRemovedConstraints, start = StartingSequence ? 1 : 0;
  b.D = D; b.m = &m; b.DPu = DPu; b.dpu_stride = (long long)(NbOfConstraints + 1) * 2 * m_CardV;
  b.DPx = DPx; b.dpx_stride = NbOfConstraints ? NbOfConstraints : 1;
  b.ZMPRef = ZMPRef; b.XkYk = XkYk; b.X = X;
  std::vector<int32_t> sim(SimilarConstraint.begin(), SimilarConstraint.end());
  b.similar = sim.empty() ? nullptr : sim.data(); b.similar_stride = (long long)sim.size();
  b.n_removed = &nrem; b.starting = &start;
  b.hot = &m_Hot; b.hot_start = 1;   // m_HotStart = true, PLDPSolver.cpp:65
  b.max_iterations = 0; b.info = &m_Info;
  check(wg_pldp_solve_batch(default_context(), WG_MEM_HOST, 1, &b), "wg_pldp_solve_batch");
  if (m_Info.status == 2) return -2;
  return m_Info.rc;
}

}  // namespace Solver
}  // namespace Optimization

namespace PatternGeneratorJRL {

// ---------------------------------------------------------------------------------------------
// ZMPRefTrajectoryGeneration (plugin commands of ZMPRefTrajectoryGeneration.cpp:49-110)
// ---------------------------------------------------------------------------------------------
ZMPRefTrajectoryGeneration::ZMPRefTrajectoryGeneration(SimplePluginManager *lSPM)
    : SimplePlugin(lSPM), m_Tsingle(0.), m_Tdble(0.), m_SamplingPeriod(0.005), m_Omega(0.), m_ComHeight(0.),
      m_StepHeight(0.), m_CurrentTime(0.), m_OnLineMode(false)
{
  std::string names[6] = {":omega", ":stepheight", ":singlesupporttime", ":doublesupporttime", ":comheight", ":samplingperiod"};
  for (auto &n : names) RegisterMethod(n);
}
void ZMPRefTrajectoryGeneration::CallMethod(std::string &Method, std::istringstream &strm)
{
  if (Method == ":omega") strm >> m_Omega;
  else if (Method == ":stepheight") strm >> m_StepHeight;
  else if (Method == ":singlesupporttime") strm >> m_Tsingle;
  else if (Method == ":doublesupporttime") strm >> m_Tdble;
  else if (Method == ":comheight") strm >> m_ComHeight;
  else if (Method == ":samplingperiod") strm >> m_SamplingPeriod;
}

// ---------------------------------------------------------------------------------------------
// ZMPVelocityReferencedQP
// ---------------------------------------------------------------------------------------------
ZMPVelocityReferencedQP::ZMPVelocityReferencedQP(SimplePluginManager *lSPM, std::string, double sole_length,
                                                 double sole_width)
    : ZMPRefTrajectoryGeneration(lSPM), m_SoleLength(sole_length), m_SoleWidth(sole_width), m_ParamsDirty(true),
      m_StepsBeforeStop(0), m_HaveLastQP(false)
{
  std::memset(&m_State, 0, sizeof m_State);
  std::memset(&m_LastQP, 0, sizeof m_LastQP);
  wg_herdt_mpc_default_params(&m_Params);
  m_Tsingle = m_Params.t_single; m_Tdble = m_Params.t_double;
  std::string names[3] = {":previewcontroltime", ":numberstepsbeforestop", ":stoppg"};
  for (auto &n : names) RegisterMethod(n);
}
ZMPVelocityReferencedQP::ZMPVelocityReferencedQP(SimplePluginManager *lSPM, std::string, CjrlHumanoidDynamicRobot *aHS)
    : ZMPRefTrajectoryGeneration(lSPM), m_SoleLength(0.25), m_SoleWidth(0.14), m_ParamsDirty(true), m_StepsBeforeStop(0),
      m_HaveLastQP(false)
{
  std::memset(&m_State, 0, sizeof m_State);
  std::memset(&m_LastQP, 0, sizeof m_LastQP);
  wg_herdt_mpc_default_params(&m_Params);
  m_Tsingle = m_Params.t_single; m_Tdble = m_Params.t_double;
  std::string names[3] = {":previewcontroltime", ":numberstepsbeforestop", ":stoppg"};
  for (auto &n : names) RegisterMethod(n);
  if (aHS) {
    // RelativeFeetInequalities::set_feet_dimensions (relative-feet-inequalities.cpp:153-182): the LEFT foot's sole wins
    double l = 0.0, w = 0.0;
    if (aHS->rightFoot()) aHS->rightFoot()->getSoleSize(l, w);
    if (aHS->leftFoot()) aHS->leftFoot()->getSoleSize(l, w);
    if (l > 0.0 && w > 0.0) { m_SoleLength = l; m_SoleWidth = w; }
    // OrientationsPreview::OrientationsPreview (OrientationsPreview.cpp:42-68)
    CjrlJoint *waist = aHS->waist();
    CjrlJoint *lh = aHS->jointsBetween(*waist, *aHS->leftFoot()->associatedAnkle())[1];
    CjrlJoint *rh = aHS->jointsBetween(*waist, *aHS->rightFoot()->associatedAnkle())[1];
    SetHipYawJoints(lh->lowerBound(0), lh->upperBound(0), rh->lowerBound(0), rh->upperBound(0), lh->upperVelocityBound(0));
  }
}
ZMPVelocityReferencedQP::~ZMPVelocityReferencedQP() {}
void ZMPVelocityReferencedQP::CallMethod(std::string &Method, std::istringstream &strm)
{
  if (Method == ":numberstepsbeforestop") {
    unsigned n = 0;
    strm >> n;
    m_StepsBeforeStop = n;
    m_State.sup_steps_left = (int32_t)n;   // CurrentSupport.NbStepsLeft + SupportFSM::NbStepsSSDS (:197-202)
    m_State.nb_steps_ssds = (int32_t)n;
  }
  if (Method == ":stoppg") m_State.ending_phase = 1;
  ZMPRefTrajectoryGeneration::CallMethod(Method, strm);
  if (Method == ":singlesupporttime" || Method == ":doublesupporttime") m_ParamsDirty = true;
}
void ZMPVelocityReferencedQP::SetHipYawJoints(double lLeft, double uLeft, double lRight, double uRight,
                                              double upperVelocityBound)
{
  const double lo = -30.0 / 180.0 * M_PI, up = 45.0 / 180.0 * M_PI;   // OrientationsPreview.cpp:50-66
  m_Params.hip_lower[0] = (lLeft == uLeft) ? lo : lLeft;   m_Params.hip_upper[0] = (lLeft == uLeft) ? up : uLeft;
  m_Params.hip_lower[1] = (lRight == uRight) ? lo : lRight; m_Params.hip_upper[1] = (lRight == uRight) ? up : uRight;
  m_Params.foot_vel_limit = fabs(upperVelocityBound);
  m_ParamsDirty = true;
}

int ZMPVelocityReferencedQP::InitOnLine(std::deque<ZMPPosition> &FinalZMPTraj_deq, std::deque<COMState> &FinalCoMPositions_deq,
                                        std::deque<FootAbsolutePosition> &FinalLeftFootTraj_deq,
                                        std::deque<FootAbsolutePosition> &FinalRightFootTraj_deq,
                                        FootAbsolutePosition &InitLeft, FootAbsolutePosition &InitRight,
                                        std::deque<RelativeFootPosition> &, COMState &lStartingCOMState,
                                        MAL_S3_VECTOR_TYPE(double) &lStartingZMPPosition)
{
  wg_ctx *ctx = default_context();
  wg_herdt_params hp;
  wg_herdt_default_params(m_SoleLength, m_SoleWidth, &hp);
  check(wg_herdt_set_params(ctx, &hp), "wg_herdt_set_params");
  if (m_Tsingle > 0.0) m_Params.t_single = m_Tsingle;
  if (m_Tdble > 0.0) m_Params.t_double = m_Tdble;
  check(wg_herdt_mpc_set_params(ctx, &m_Params), "wg_herdt_mpc_set_params");
  m_ParamsDirty = false;
  const double init9[9] = {lStartingCOMState.x[0], lStartingCOMState.y[0], lStartingCOMState.z[0],
                           InitLeft.x, InitLeft.y, InitLeft.theta, InitRight.x, InitRight.y, InitRight.theta};
  const double keep_ref[3] = {m_State.new_ref[0], m_State.new_ref[1], m_State.new_ref[2]};
  const int32_t keep_stop = m_State.ending_phase;
  check(wg_herdt_mpc_init(ctx, WG_MEM_HOST, 1, init9, 0, &m_State), "wg_herdt_mpc_init");
  for (int i = 0; i < 3; ++i) m_State.new_ref[i] = keep_ref[i];
  m_State.ending_phase = keep_stop;
  if (m_StepsBeforeStop) { m_State.sup_steps_left = (int32_t)m_StepsBeforeStop; m_State.nb_steps_ssds = (int32_t)m_StepsBeforeStop; }
  // CoM_ takes the whole starting state and OrientPrw_->CurrentTrunkState its yaw (ZMPVelocityReferencedQP.cpp:283-301):
  // wg_herdt_mpc_init only knows the position, the derivatives are written into the (host-editable) state record here
  for (int i = 1; i < 3; ++i) {
    m_State.com_x[i] = lStartingCOMState.x[i]; m_State.com_y[i] = lStartingCOMState.y[i];
    m_State.com_front[i] = lStartingCOMState.x[i]; m_State.com_front[3 + i] = lStartingCOMState.y[i];
    m_State.com_back[i] = lStartingCOMState.x[i]; m_State.com_back[3 + i] = lStartingCOMState.y[i];
  }
  for (int i = 0; i < 3; ++i) m_State.trunk_yaw[i] = lStartingCOMState.yaw[i];
  m_State.com_back[7] = lStartingCOMState.yaw[0]; m_State.com_back[8] = lStartingCOMState.yaw[1];
  m_HaveLastQP = false;
  // the TimeBuffer_/m_SamplingPeriod buffered start samples (ZMPVelocityReferencedQP.cpp:243-271)
  const int AddArraySize = (int)(m_Params.time_buffer / m_Params.Ts);
  FinalZMPTraj_deq.assign(AddArraySize, ZMPPosition());
  FinalCoMPositions_deq.assign(AddArraySize, COMState());
  FinalLeftFootTraj_deq.assign(AddArraySize, InitLeft);
  FinalRightFootTraj_deq.assign(AddArraySize, InitRight);
  double t = 0.0;
  for (int i = 0; i < AddArraySize; ++i) {
    ZMPPosition &z = FinalZMPTraj_deq[i];
    z.px = lStartingZMPPosition[0]; z.py = lStartingZMPPosition[1]; z.pz = lStartingZMPPosition[2];
    z.theta = 0.0; z.time = t; z.stepType = 0;
    FinalCoMPositions_deq[i] = lStartingCOMState;
    FinalLeftFootTraj_deq[i].time = FinalRightFootTraj_deq[i].time = t;
    FinalLeftFootTraj_deq[i].stepType = FinalRightFootTraj_deq[i].stepType = 10;
    t += m_Params.Ts;
  }
  m_State.com_back[9] = lStartingZMPPosition[0];
  m_State.com_back[10] = lStartingZMPPosition[1];
  m_OnLineMode = true;
  return 0;
}

solution_t &ZMPVelocityReferencedQP::Solution()
{
  if (!m_HaveLastQP) return m_Solution;
  wg_herdt_qp_output out;
  check(wg_herdt_qp_solve_batch(default_context(), WG_MEM_HOST, 1, &m_LastQP, &out), "wg_herdt_qp_solve_batch");
  solution_t &S = m_Solution;
  S.NbVariables = (unsigned)out.n_vars; S.NbConstraints = (unsigned)out.n_rows; S.Fail = out.fail;
  S.Solution_vec.assign(out.x, out.x + out.n_vars);
  S.ConstrLagr_vec.assign(out.lagr, out.lagr + out.n_rows);
  S.LBoundsLagr_vec.assign(out.n_vars, 0.0); S.UBoundsLagr_vec.assign(out.n_vars, 0.0);   // the +-1e8 box never binds
  S.SupportStates_deq.clear(); S.SupportOrientations_deq.clear(); S.TrunkOrientations_deq.clear();
  for (int i = 0; i <= WG_HERDT_N; ++i) {
    support_state_t st;
    st.Phase = m_LastQP.sup_phase[i]; st.Foot = m_LastQP.sup_foot[i]; st.StepNumber = (unsigned)m_LastQP.sup_step[i];
    st.StateChanged = m_LastQP.sup_changed[i] != 0;
    st.X = m_LastQP.sup_x[i]; st.Y = m_LastQP.sup_y[i]; st.Yaw = m_LastQP.sup_yaw[i];
    S.SupportStates_deq.push_back(st);
    if (i > 0 && st.StateChanged && st.StepNumber > 0) S.SupportOrientations_deq.push_back(st.Yaw);
  }
  m_HaveLastQP = false;
  return S;
}

static void tick_to_rows(const double *com11, const wg_herdt_foot_sample &L, const wg_herdt_foot_sample &R, double time,
                         COMState &c, ZMPPosition &z, FootAbsolutePosition &lf, FootAbsolutePosition &rf)
{
  c.reset();
  for (int i = 0; i < 3; ++i) { c.x[i] = com11[i]; c.y[i] = com11[3 + i]; }
  c.z[0] = com11[6]; c.yaw[0] = com11[7]; c.yaw[1] = com11[8];
  std::memset(&z, 0, sizeof z);
  z.px = com11[9]; z.py = com11[10]; z.time = time;
  const wg_herdt_foot_sample *src[2] = {&L, &R};
  FootAbsolutePosition *dst[2] = {&lf, &rf};
  for (int f = 0; f < 2; ++f) {
    FootAbsolutePosition &o = *dst[f];
    std::memset(&o, 0, sizeof o);
    o.x = src[f]->x; o.y = src[f]->y; o.z = src[f]->z; o.theta = src[f]->theta;
    o.dx = src[f]->dx; o.dy = src[f]->dy; o.dz = src[f]->dz; o.dtheta = src[f]->dtheta;
    o.ddx = src[f]->ddx; o.ddy = src[f]->ddy;
    o.time = time;
  }
}

void ZMPVelocityReferencedQP::OnLine(double time, std::deque<ZMPPosition> &FinalZMPTraj_deq,
                                     std::deque<COMState> &FinalCOMTraj_deq,
                                     std::deque<FootAbsolutePosition> &FinalLeftFootTraj_deq,
                                     std::deque<FootAbsolutePosition> &FinalRightFootTraj_deq)
{
  // ZMPVelocityReferencedQP.cpp:331-346.  On the ticks where no QP fires only the end-of-online-mode test runs (here);
  // on a firing tick the device loop runs both tests itself with the same clock value.
  if (!m_State.online_mode) { m_OnLineMode = false; return; }
  if (!(time + 0.00001 > m_State.upper_time_limit)) {
    if (m_State.ending_phase && time >= m_State.time_to_stop) m_State.online_mode = 0;
    m_OnLineMode = m_State.online_mode != 0;
    return;
  }
  wg_ctx *ctx = default_context();
  if (m_ParamsDirty) {
    if (m_Tsingle > 0.0) m_Params.t_single = m_Tsingle;
    if (m_Tdble > 0.0) m_Params.t_double = m_Tdble;
    check(wg_herdt_mpc_set_params(ctx, &m_Params), "wg_herdt_mpc_set_params");
    m_ParamsDirty = false;
  }
  // the caller owns the clock: make the device loop's first tick land exactly on it (clock + Ts == time)
  double c0 = time - m_Params.Ts;
  for (int guard = 0; guard < 4 && c0 + m_Params.Ts != time; ++guard)
    c0 = std::nextafter(c0, (c0 + m_Params.Ts < time) ? 1e300 : -1e300);
  m_State.clock = c0;
  wg_herdt_tick rows[WG_HERDT_TICKS_PER_STEP];
  std::memset(rows, 0, sizeof rows);
  check(wg_herdt_mpc_run_batch(ctx, WG_MEM_HOST, 1, 1, &m_State, nullptr, rows, nullptr, &m_LastQP), "wg_herdt_mpc_run_batch");
  m_HaveLastQP = true;
  m_OnLineMode = m_State.online_mode != 0;
  // deques: the inherited last element is final now (the feet part may have been rewritten), then 19 new samples,
  // then the 20th, which the state keeps as the not-yet-final element
  COMState c; ZMPPosition z; FootAbsolutePosition lf, rf;
  const int n = WG_HERDT_TICKS_PER_STEP;
  tick_to_rows(reinterpret_cast<const double *>(&rows[0]), rows[0].left, rows[0].right, time, c, z, lf, rf);
  if (!FinalLeftFootTraj_deq.empty()) { lf.time = FinalLeftFootTraj_deq.back().time; FinalLeftFootTraj_deq.back() = lf; }
  if (!FinalRightFootTraj_deq.empty()) { rf.time = FinalRightFootTraj_deq.back().time; FinalRightFootTraj_deq.back() = rf; }
  for (int k = 1; k <= n; ++k) {
    if (k < n)
      tick_to_rows(reinterpret_cast<const double *>(&rows[k]), rows[k].left, rows[k].right, time + k * m_Params.Ts, c, z, lf, rf);
    else
      tick_to_rows(m_State.com_back, m_State.foot[0][2], m_State.foot[1][2], time + k * m_Params.Ts, c, z, lf, rf);
    FinalCOMTraj_deq.push_back(c);
    FinalZMPTraj_deq.push_back(z);
    FinalLeftFootTraj_deq.push_back(lf);
    FinalRightFootTraj_deq.push_back(rf);
  }
}

// ---------------------------------------------------------------------------------------------
// PatternGeneratorInterface (Herdt path)
// ---------------------------------------------------------------------------------------------
// (PatternGeneratorInterface: see walkgen_host_pgi.cpp)

}  // namespace PatternGeneratorJRL

// ---- Dimitrov2008 path -------------------------------------------------------------------------------------------
namespace PatternGeneratorJRL {

using walkgen_b200::walkgen_b200_check;

void ComputeConvexHull::DoComputeConvexHull(std::vector<CH_Point> aVecOfPoints, std::vector<CH_Point> &TheConvexHull)
{
  if (aVecOfPoints.empty()) return;
  if (aVecOfPoints.size() > 8) throw std::runtime_error("walkgen_b200: DoComputeConvexHull is built for at most 8 points (two feet)");
  wg_ctx *ctx = walkgen_b200::default_context();
  double xy[16], hull[16];
  int32_t n = 0;
  for (size_t i = 0; i < aVecOfPoints.size(); ++i) { xy[2 * i] = aVecOfPoints[i].col; xy[2 * i + 1] = aVecOfPoints[i].row; }
  walkgen_b200_check(wg_convex_hull_batch(ctx, WG_MEM_HOST, 1, (int)aVecOfPoints.size(), xy, hull, &n), "wg_convex_hull_batch");
  for (int i = 0; i < n; ++i) { CH_Point p; p.col = hull[2 * i]; p.row = hull[2 * i + 1]; TheConvexHull.push_back(p); }
}

FootConstraintsAsLinearSystem::FootConstraintsAsLinearSystem(SimplePluginManager *aSPM, double sole_length, double sole_width)
    : SimplePlugin(aSPM), m_SoleLength(sole_length), m_SoleWidth(sole_width) {}

int FootConstraintsAsLinearSystem::BuildLinearConstraintInequalities(
    std::deque<FootAbsolutePosition> &Left, std::deque<FootAbsolutePosition> &Right,
    std::deque<LinearConstraintInequality_t *> &Queue, double ConstraintOnX, double ConstraintOnY)
{
  if (Left.size() != Right.size()) return -1;
  const size_t n = Left.size();
  if (n == 0) return 0;
  wg_ctx *ctx = walkgen_b200::default_context();
  wg_dimitrov_params p;
  wg_dimitrov_default_params(&p);
  p.constraint_x = ConstraintOnX; p.constraint_y = ConstraintOnY;
  p.sole_length = m_SoleLength; p.sole_width = m_SoleWidth;
  walkgen_b200_check(wg_dimitrov_set_params(ctx, &p, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr), "wg_dimitrov_set_params");
  std::vector<wg_foot_sample> l(n), r(n);
  std::vector<int32_t> ty(3 * n, 0);
  for (size_t i = 0; i < n; ++i) {
    l[i].x = Left[i].x; l[i].y = Left[i].y; l[i].z = Left[i].z; l[i].theta = Left[i].theta; l[i].omega = l[i].omega2 = 0;
    r[i].x = Right[i].x; r[i].y = Right[i].y; r[i].z = Right[i].z; r[i].theta = Right[i].theta; r[i].omega = r[i].omega2 = 0;
    ty[3 * i + 1] = Left[i].stepType; ty[3 * i + 2] = Right[i].stepType;
  }
  // a state change needs at least one sample: n polygons is the hard bound, a walk has a few per step
  const int64_t cap = (int64_t)std::min<size_t>(n, 4096);
  const int64_t so[2] = {0, (int64_t)n}, lo[2] = {0, cap};
  std::vector<wg_lci> out((size_t)cap);
  int32_t cnt = 0;
  walkgen_b200_check(wg_fcals_build_batch(ctx, WG_MEM_HOST, 1, so, l.data(), r.data(), ty.data(), lo, out.data(), &cnt),
                     "wg_fcals_build_batch");
  for (int k = 0; k < cnt; ++k) {
    LinearConstraintInequality_t *q = new LinearConstraintInequality_t;
    const int nr = out[k].rows;
    q->A.resize(nr, 2); q->B.resize(nr, 1);
    q->Center.assign(out[k].center, out[k].center + 2);
    q->SimilarConstraints.assign(out[k].similar, out[k].similar + nr);
    for (int j = 0; j < nr; ++j) { q->A(j, 0) = out[k].A[j][0]; q->A(j, 1) = out[k].A[j][1]; q->B(j, 0) = out[k].B[j]; }
    // the clock of the feet samples as the caller set it (the device uses the accumulated 5 ms clock, which is what
    // ZMPDiscretization writes into .time)
    q->StartingTime = Left[out[k].first_sample].time;
    q->EndingTime = (k + 1 < cnt) ? Left[out[k + 1].first_sample].time : Left[n - 1].time;
    Queue.push_back(q);
  }
  return 0;
}

ZMPConstrainedQPFastFormulation::ZMPConstrainedQPFastFormulation(SimplePluginManager *lSPM, std::string, double sole_length,
                                                                 double sole_width)
    : ZMPRefTrajectoryGeneration(lSPM), m_Dirty(true), m_Status(0), m_Done(0)
{
  wg_dimitrov_default_params(&m_Par);
  m_Par.sole_length = sole_length; m_Par.sole_width = sole_width;
  wg_zmpdisc_default_params(&m_Zd);
  std::string name = ":setdimitrovconstraint";
  RegisterMethod(name);
}

int ZMPConstrainedQPFastFormulation::InitConstants()
{
  wg_ctx *ctx = walkgen_b200::default_context();
  const int rc = wg_dimitrov_set_params(ctx, &m_Par, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc == WG_OK) m_Dirty = false;
  return rc == WG_OK ? 0 : -1;
}

void ZMPConstrainedQPFastFormulation::CallMethod(std::string &Method, std::istringstream &strm)
{
  if (Method == ":setdimitrovconstraint") {       /* ZMPConstrainedQPFastFormulation.cpp:1586-1596 */
    strm >> m_Par.constraint_x >> m_Par.constraint_y;
    m_Dirty = true;
    return;
  }
  ZMPRefTrajectoryGeneration::CallMethod(Method, strm);
}

void ZMPConstrainedQPFastFormulation::GetZMPDiscretization(
    std::deque<ZMPPosition> &ZMPPositions, std::deque<COMState> &COMStates, std::deque<RelativeFootPosition> &Rel,
    std::deque<FootAbsolutePosition> &Left, std::deque<FootAbsolutePosition> &Right, double, COMState &,
    MAL_S3_VECTOR_TYPE(double) &, FootAbsolutePosition &InitLeft, FootAbsolutePosition &InitRight)
{
  if (Rel.empty()) return;
  wg_ctx *ctx = walkgen_b200::default_context();
  m_Zd.sampling_period = m_SamplingPeriod; m_Zd.t_single = m_Tsingle; m_Zd.t_double = m_Tdble;
  m_Zd.step_height = m_StepHeight; m_Zd.omega = m_Omega;
  m_Par.sampling_period = m_SamplingPeriod;
  // the context's constants are shared by every generator object: (re)install this object's before each plan
  if (InitConstants() != 0) throw std::runtime_error(std::string("walkgen_b200: ") + wg_last_error(ctx));
  std::vector<wg_rel_step> steps(Rel.size());
  for (size_t i = 0; i < Rel.size(); ++i) {
    std::memset(&steps[i], 0, sizeof(wg_rel_step));
    steps[i].sx = Rel[i].sx; steps[i].sy = Rel[i].sy; steps[i].theta = Rel[i].theta;
    steps[i].ss_time = Rel[i].SStime; steps[i].ds_time = Rel[i].DStime; steps[i].step_type = Rel[i].stepType;
  }
  const int64_t off[2] = {0, (int64_t)steps.size()};
  const double feet[6] = {InitLeft.x, InitLeft.y, InitLeft.theta, InitRight.x, InitRight.y, InitRight.theta};
  wg_kajita_plan *plan = nullptr;
  walkgen_b200_check(wg_kajita_plan_create(ctx, &m_Zd, 1, off, steps.data(), feet, &plan), "wg_kajita_plan_create");
  const int64_t n = wg_kajita_plan_total_samples(plan);
  std::vector<double> com(6 * (size_t)n), zmp(2 * (size_t)n);
  std::vector<wg_foot_sample> l((size_t)n), r((size_t)n);
  int32_t status = 0, done = 0;
  const int rc = wg_dimitrov_run_batch(ctx, plan, WG_MEM_HOST, com.data(), zmp.data(), l.data(), r.data(), nullptr, nullptr,
                                       &status, &done);
  wg_kajita_plan_destroy(plan);
  walkgen_b200_check(rc, "wg_dimitrov_run_batch");
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
