// walkgen_host_pgi.cpp - host-side mirror of the reference's step stack, ZMPDiscretization and PatternGeneratorInterface
// (walkgen_host.hh) over the C ABI.  No algorithm lives here: footsteps -> ZMP reference / feet is zmpdisc_kernel
// (wg_zmpdisc_run_batch), the CoM is preview_fused_kernel (wg_preview_run_batch), the Herdt generator is the closed-loop
// kernels behind wg_herdt_mpc_run_batch.
#include "walkgen_host.hh"
#include <cmath>
#include <cstdlib>
#include <iostream>

using walkgen_b200::default_context;

namespace walkgen_b200 { void walkgen_b200_check(int rc, const char *what); }
using walkgen_b200::walkgen_b200_check;

namespace PatternGeneratorJRL {

// ---------------------------------------------------------------------------------------------------------------------
// StepStackHandler (src/StepStackHandler.cpp)
// ---------------------------------------------------------------------------------------------------------------------
StepStackHandler::StepStackHandler(SimplePluginManager *lSPM)
    : SimplePlugin(lSPM), m_SingleSupportTime(0.78), m_DoubleSupportTime(0.02), m_WalkMode(0),
      m_KeepLastCorrectSupportFoot(1), m_OnLineSteps(false), m_TransitionFinishOnLine(false)
{
  // the reference registers 7 of its 8 names (loop bound 7, StepStackHandler.cpp:64): ":arccentered" stays unregistered
  std::string names[7] = {":walkmode", ":singlesupporttime", ":doublesupporttime", ":supportfoot", ":lastsupport", ":arc",
                          ":addstandardonlinestep"};
  for (auto &n : names) RegisterMethod(n);
}

static RelativeFootPosition make_step(double sx, double sy, double theta, double ss, double ds, int type)
{
  RelativeFootPosition f;
  f.sx = sx; f.sy = sy; f.theta = theta; f.SStime = ss; f.DStime = ds; f.stepType = type; f.DeviationHipHeight = 0.0;
  return f;
}

void StepStackHandler::ReadStepSequenceAccordingToWalkMode(std::istringstream &strm)
{
  m_RelativeFootPositions.clear();
  const bool hip = (m_WalkMode == 1 || m_WalkMode == 3), timed = (m_WalkMode == 5);
  if (!(m_WalkMode == 0 || m_WalkMode == 4 || hip || timed)) return;   // mode 2 (step-over planner) is outside this path
  while (!strm.eof()) {
    RelativeFootPosition f = make_step(0, 0, 0, m_SingleSupportTime, m_DoubleSupportTime, 1);
    if (strm.eof()) break;
    strm >> f.sx;
    if (strm.eof()) break;
    strm >> f.sy;
    if (strm.eof()) break;
    strm >> f.theta;
    if (hip) { if (strm.eof()) break; strm >> f.DeviationHipHeight; }
    if (timed) {
      if (strm.eof()) break;
      strm >> f.SStime;
      if (strm.eof()) break;
      strm >> f.DStime;
    }
    if (strm.fail()) break;
    m_RelativeFootPositions.push_back(f);
    m_KeepLastCorrectSupportFoot = (f.sy > 0) ? -1 : 1;
  }
}
void StepStackHandler::m_PartialStepSequence(std::istringstream &strm)
{
  while (!strm.eof()) {
    RelativeFootPosition f = make_step(0, 0, 0, m_SingleSupportTime, m_DoubleSupportTime, 0);
    if (strm.eof()) break;
    strm >> f.sx;
    if (strm.eof()) break;
    strm >> f.sy;
    if (strm.eof()) break;
    strm >> f.theta;
    if (strm.fail()) break;
    m_RelativeFootPositions.push_back(f);
  }
}
void StepStackHandler::CreateArcInStepStack(double x, double y, double, double arc_deg, int SupportFoot)
{
  // host arithmetic of the C ABI (wg_steps_arc restates StepStackHandler.cpp:299-457; pinned by the Circle datref)
  std::vector<wg_rel_step> st(256);
  int n = 0, keep = m_KeepLastCorrectSupportFoot;
  walkgen_b200_check(wg_steps_arc(st.data(), (int)st.size(), &n, x, y, arc_deg, SupportFoot, m_SingleSupportTime,
                                  m_DoubleSupportTime, &keep), "wg_steps_arc");
  for (int i = 0; i < n; ++i)
    m_RelativeFootPositions.push_back(make_step(st[i].sx, st[i].sy, st[i].theta, st[i].ss_time, st[i].ds_time, st[i].step_type));
  m_KeepLastCorrectSupportFoot = keep;
}
void StepStackHandler::CreateArcCenteredInStepStack(double, double, int)
{
  throw std::runtime_error("walkgen_b200: CreateArcCenteredInStepStack is not part of the accelerated path (:arccentered is "
                           "never registered by the reference either)");
}
void StepStackHandler::PrepareForSupportFoot(int SupportFoot)
{
  m_RelativeFootPositions.push_back(make_step(0, SupportFoot * 0.095, 0, m_SingleSupportTime, m_DoubleSupportTime, 0));
}
void StepStackHandler::FinishOnTheLastCorrectSupportFoot()
{
  m_RelativeFootPositions.push_back(make_step(0, m_KeepLastCorrectSupportFoot * 0.19, 0, m_SingleSupportTime,
                                              m_DoubleSupportTime, 0));
}
void StepStackHandler::AddStepInTheStack(double sx, double sy, double theta, double sstime, double dstime)
{
  m_RelativeFootPositions.push_back(make_step(sx, sy, theta, sstime, dstime, 0));
}
void StepStackHandler::AddStandardOnLineStep(bool NewStep, double NewStepX, double NewStepY, double NewTheta)
{
  if (!m_OnLineSteps) return;
  if (!NewStep)
    m_RelativeFootPositions.push_back(make_step(0, m_KeepLastCorrectSupportFoot * 0.19, 0, m_SingleSupportTime, m_DoubleSupportTime, 0));
  else
    m_RelativeFootPositions.push_back(make_step(NewStepX, NewStepY + m_KeepLastCorrectSupportFoot * 0.19, NewTheta,
                                                m_SingleSupportTime, m_DoubleSupportTime, 0));
  m_KeepLastCorrectSupportFoot = -m_KeepLastCorrectSupportFoot;
}
bool StepStackHandler::RemoveFirstStepInTheStack()
{
  if (!m_RelativeFootPositions.empty()) m_RelativeFootPositions.pop_front();
  if (m_RelativeFootPositions.empty() && m_TransitionFinishOnLine && m_OnLineSteps) {
    m_OnLineSteps = false;
    m_TransitionFinishOnLine = false;
    return true;
  }
  return false;
}
void StepStackHandler::StopOnLineStep()
{
  m_TransitionFinishOnLine = true;
  if (m_RelativeFootPositions.size() % 2 == 0) m_KeepLastCorrectSupportFoot = -m_KeepLastCorrectSupportFoot;
  m_RelativeFootPositions.clear();
}
void StepStackHandler::CopyRelativeFootPosition(std::deque<RelativeFootPosition> &l, bool PerformClean)
{
  l.assign(m_RelativeFootPositions.begin(), m_RelativeFootPositions.end());
  if (PerformClean) m_RelativeFootPositions.clear();
}
bool StepStackHandler::ReturnFrontFootPosition(RelativeFootPosition &aRFP)
{
  if (m_RelativeFootPositions.empty()) return false;
  aRFP = m_RelativeFootPositions.front();
  return true;
}
void StepStackHandler::CallMethod(std::string &Method, std::istringstream &strm)
{
  if (Method == ":singlesupporttime") strm >> m_SingleSupportTime;
  else if (Method == ":doublesupporttime") strm >> m_DoubleSupportTime;
  else if (Method == ":walkmode") strm >> m_WalkMode;
  else if (Method == ":supportfoot") { int f = -1; strm >> f; PrepareForSupportFoot(f); }
  else if (Method == ":lastsupport") FinishOnTheLastCorrectSupportFoot();
  else if (Method == ":addstandardonlinestep") {
    double x = 0, y = 0, th = 0;
    while (!strm.eof()) { strm >> x; if (strm.eof()) break; strm >> y; if (strm.eof()) break; strm >> th; if (strm.fail()) break; }
    AddStandardOnLineStep(true, x, y, th);
  } else if (Method == ":arc") {
    double x = 0, y = 0, arc = 0;
    int foot = -1;
    while (!strm.eof()) {
      strm >> x; if (strm.eof()) break;
      strm >> y; if (strm.eof()) break;
      strm >> arc; if (strm.eof()) break;
      strm >> foot; if (strm.fail()) break;
    }
    CreateArcInStepStack(x, y, 0.0, arc, foot);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// ZMPDiscretization (src/ZMPRefTrajectoryGeneration/ZMPDiscretization.cpp)
// ---------------------------------------------------------------------------------------------------------------------
ZMPDiscretization::ZMPDiscretization(SimplePluginManager *lSPM, std::string, CjrlHumanoidDynamicRobot *aHS)
    : ZMPRefTrajectoryGeneration(lSPM), m_PreviewControlTime(1.6), m_Emitted(0)
{
  wg_zmpdisc_default_params(&m_Zd);
  m_Tsingle = m_Zd.t_single; m_Tdble = m_Zd.t_double; m_SamplingPeriod = m_Zd.sampling_period;
  m_StepHeight = m_Zd.step_height; m_Omega = m_Zd.omega; m_ComHeight = 0.814;
  for (double &v : m_InitFeet) v = 0.0;
  if (aHS && aHS->leftFoot()) {
    // m_FootB / m_FootH / m_FootF from the sole and the ankle position (ZMPDiscretization.cpp:59-80 through
    // FootTrajectoryGenerationStandard::InitializeInternalDataStructures)
    double l = 0.0, w = 0.0;
    vector3d ankle;
    aHS->leftFoot()->getSoleSize(l, w);
    aHS->leftFoot()->getAnklePositionInLocalFrame(ankle);
    if (l > 0.0) { m_Zd.foot_b = 0.5 * l + ankle[0]; m_Zd.foot_f = 0.5 * l - ankle[0]; m_Zd.foot_h = ankle[2]; }
  }
  std::string names[3] = {":prevzmpinitprofil", ":zeroinitprofil", ":previewcontroltime"};   // ZMPDiscretization.cpp:1304-1307
  for (auto &n : names) RegisterMethod(n);
}
ZMPDiscretization::~ZMPDiscretization() {}
void ZMPDiscretization::CallMethod(std::string &Method, std::istringstream &strm)
{
  if (Method == ":previewcontroltime") { strm >> m_PreviewControlTime; return; }
  if (Method == ":prevzmpinitprofil" || Method == ":zeroinitprofil") return;   // start-profile switches: the kernel implements
                                                                               // the default (zero) start profile
  ZMPRefTrajectoryGeneration::CallMethod(Method, strm);
}
void ZMPDiscretization::SetZMPShift(std::vector<double> &ZMPShift)
{
  for (size_t i = 0; i < 4 && i < ZMPShift.size(); ++i) m_Zd.zmp_shift[i] = ZMPShift[i];
}
int ZMPDiscretization::ReturnOptimalTimeToRegenerateAStep()
{
  return 2 * (int)(m_PreviewControlTime / m_SamplingPeriod);   // ZMPDiscretization.cpp:1122-1127
}

void ZMPDiscretization::emit(std::deque<ZMPPosition> &Z, std::deque<COMState> &Cq, std::deque<FootAbsolutePosition> &L,
                             std::deque<FootAbsolutePosition> &R, bool with_end_phase)
{
  if (m_Steps.empty()) return;
  wg_ctx *ctx = default_context();
  m_Zd.sampling_period = m_SamplingPeriod; m_Zd.preview_time = m_PreviewControlTime;
  m_Zd.t_single = m_Tsingle; m_Zd.t_double = m_Tdble; m_Zd.step_height = m_StepHeight; m_Zd.omega = m_Omega;
  const int64_t off[2] = {0, (int64_t)m_Steps.size()};
  wg_kajita_plan *plan = nullptr;
  walkgen_b200_check(wg_kajita_plan_create(ctx, &m_Zd, 1, off, m_Steps.data(), m_InitFeet, &plan), "wg_kajita_plan_create");
  const int64_t n = wg_kajita_plan_total_samples(plan);
  std::vector<double> zmp(2 * (size_t)n), th((size_t)n);
  std::vector<wg_foot_sample> l((size_t)n), r((size_t)n);
  std::vector<int32_t> ty(3 * (size_t)n);
  const int rc = wg_zmpdisc_run_batch(ctx, plan, WG_MEM_HOST, zmp.data(), th.data(), l.data(), r.data(), ty.data());
  wg_kajita_plan_destroy(plan);
  walkgen_b200_check(rc, "wg_zmpdisc_run_batch");
  // the end phase is the last segment: round(Tdble / (2 T)) + 3 NL samples (ZMPDiscretization.cpp:1147-1148, :1239-1241)
  const int64_t NL = (int64_t)(m_PreviewControlTime / m_SamplingPeriod);
  const int64_t tail = (int64_t)std::llround(m_Tdble / (2.0 * m_SamplingPeriod)) + 3 * NL;
  const int64_t upto = with_end_phase ? n : n - tail;
  double t = 0.0;
  for (int64_t i = 0; i < upto; ++i, t += m_SamplingPeriod) {   // the reference's clock: the period accumulated sample by sample
    if (i < m_Emitted) continue;
    ZMPPosition z;
    z.px = zmp[2 * i]; z.py = zmp[2 * i + 1]; z.pz = 0.0; z.theta = th[(size_t)i]; z.time = t; z.stepType = ty[3 * i];
    Z.push_back(z);
    COMState c;                         // ZMPDiscretization.cpp:976-1000: height and yaw only, the preview fills the rest
    c.z[0] = m_ComHeight; c.yaw[0] = z.theta;
    Cq.push_back(c);
    const wg_foot_sample *fs[2] = {&l[(size_t)i], &r[(size_t)i]};
    std::deque<FootAbsolutePosition> *dq[2] = {&L, &R};
    for (int f = 0; f < 2; ++f) {
      FootAbsolutePosition a;
      std::memset(&a, 0, sizeof a);
      a.x = fs[f]->x; a.y = fs[f]->y; a.z = fs[f]->z; a.theta = fs[f]->theta; a.omega = fs[f]->omega; a.omega2 = fs[f]->omega2;
      a.time = t; a.stepType = ty[3 * i + 1 + f];
      dq[f]->push_back(a);
    }
  }
  if (upto > m_Emitted) m_Emitted = upto;
}

static wg_rel_step to_rel(const RelativeFootPosition &f)
{
  wg_rel_step s;
  std::memset(&s, 0, sizeof s);
  s.sx = f.sx; s.sy = f.sy; s.theta = f.theta; s.ss_time = f.SStime; s.ds_time = f.DStime;
  s.step_type = f.stepType ? f.stepType : 1;
  return s;
}

int ZMPDiscretization::InitOnLine(std::deque<ZMPPosition> &Z, std::deque<COMState> &Cq, std::deque<FootAbsolutePosition> &L,
                                  std::deque<FootAbsolutePosition> &R, FootAbsolutePosition &InitLeft,
                                  FootAbsolutePosition &InitRight, std::deque<RelativeFootPosition> &Rel, COMState &,
                                  MAL_S3_VECTOR_TYPE(double) &)
{
  // ZMPDiscretization.cpp:319-513: the lead-in (2 NL samples) and then OnLineAddFoot for every step of the stack but the
  // first (which only says where the first support foot is)
  m_Steps.clear();
  m_Emitted = 0;
  m_InitFeet[0] = InitLeft.x; m_InitFeet[1] = InitLeft.y; m_InitFeet[2] = InitLeft.theta;
  m_InitFeet[3] = InitRight.x; m_InitFeet[4] = InitRight.y; m_InitFeet[5] = InitRight.theta;
  for (size_t i = 0; i < Rel.size(); ++i) m_Steps.push_back(to_rel(Rel[i]));
  if (m_Steps.empty()) return 0;
  emit(Z, Cq, L, R, false);
  m_OnLineMode = true;
  return (int)Rel.size();
}
void ZMPDiscretization::OnLineAddFoot(RelativeFootPosition &New, std::deque<ZMPPosition> &Z, std::deque<COMState> &Cq,
                                      std::deque<FootAbsolutePosition> &L, std::deque<FootAbsolutePosition> &R, bool EndSequence)
{
  m_Steps.push_back(to_rel(New));
  emit(Z, Cq, L, R, false);
  if (EndSequence) EndPhaseOfTheWalking(Z, Cq, L, R);   // ZMPDiscretization.cpp:1011-1019
}
void ZMPDiscretization::EndPhaseOfTheWalking(std::deque<ZMPPosition> &Z, std::deque<COMState> &Cq,
                                             std::deque<FootAbsolutePosition> &L, std::deque<FootAbsolutePosition> &R)
{
  emit(Z, Cq, L, R, true);
  m_OnLineMode = false;
}
void ZMPDiscretization::GetZMPDiscretization(std::deque<ZMPPosition> &Z, std::deque<COMState> &Cq,
                                             std::deque<RelativeFootPosition> &Rel, std::deque<FootAbsolutePosition> &L,
                                             std::deque<FootAbsolutePosition> &R, double, COMState &lStartingCOMState,
                                             MAL_S3_VECTOR_TYPE(double) &lStartingZMPPosition,
                                             FootAbsolutePosition &InitLeft, FootAbsolutePosition &InitRight)
{
  // ZMPDiscretization.cpp:142-175: InitOnLine + EndPhaseOfTheWalking (one pass of the kernel here)
  m_Steps.clear();
  m_Emitted = 0;
  m_InitFeet[0] = InitLeft.x; m_InitFeet[1] = InitLeft.y; m_InitFeet[2] = InitLeft.theta;
  m_InitFeet[3] = InitRight.x; m_InitFeet[4] = InitRight.y; m_InitFeet[5] = InitRight.theta;
  for (size_t i = 0; i < Rel.size(); ++i) m_Steps.push_back(to_rel(Rel[i]));
  (void)lStartingCOMState; (void)lStartingZMPPosition;
  emit(Z, Cq, L, R, true);
  Cq.resize(Z.size());
}

// ---------------------------------------------------------------------------------------------------------------------
// PatternGeneratorInterface
// ---------------------------------------------------------------------------------------------------------------------
void PatternGeneratorInterface::construct(double sole_length, double sole_width, CjrlHumanoidDynamicRobot *aHDR)
{
  m_InternalClock = 0.0; m_AlgorithmforZMPCOM = 0; m_Running = false;
  m_AutoFirstStep = false; m_KajitaOnLine = false; m_PreviewedUpTo = 0;
  for (double &v : m_PreviewState) v = 0.0;
  m_ZMPShift.assign(4, 0.0);
  m_PC = new PreviewControl(this, OptimalControllerSolver::MODE_WITHOUT_INITIALPOS, true);   // PGI.cpp:254
  if (aHDR) m_ZMPVRQP = new ZMPVelocityReferencedQP(this, "", aHDR);                         // PGI.cpp:247
  else m_ZMPVRQP = new ZMPVelocityReferencedQP(this, "", sole_length, sole_width);
  m_StepStackHandler = new StepStackHandler(this);                                           // PGI.cpp:209
  m_ZMPD = new ZMPDiscretization(this, "", aHDR);                                            // PGI.cpp:235
  // start configuration of the reference's sample robot in half-sitting (TestHerdt2010 datref, line 1)
  m_StartCOM.x[0] = 0.0316055; m_StartCOM.y[0] = 0.0; m_StartCOM.z[0] = 0.7116911;
  std::memset(&m_StartLF, 0, sizeof m_StartLF); std::memset(&m_StartRF, 0, sizeof m_StartRF);
  m_StartLF.y = 0.09; m_StartRF.y = -0.09;
  // PGI.cpp:186-201
  std::string names[] = {":LimitsFeasibility", ":ZMPShiftParameters", ":TimeDistributionParameters", ":stepseq", ":finish",
                         ":StartOnLineStepSequencing", ":StopOnLineStepSequencing", ":readfilefromkw",
                         ":SetAlgoForZmpTrajectory", ":SetAutoFirstStep", ":ChangeNextStep", ":samplingperiod",
                         ":HerdtOnline", ":setVelReference", ":setCoMPerturbationForce"};
  for (auto &n : names) SimplePlugin::RegisterMethod(n);
}
PatternGeneratorInterface::PatternGeneratorInterface(double sole_length, double sole_width)
    : SimplePlugin(this), m_OwnRobot(nullptr)
{
  construct(sole_length, sole_width, nullptr);
}
PatternGeneratorInterface::PatternGeneratorInterface(CjrlHumanoidDynamicRobot *aHDR) : SimplePlugin(this), m_OwnRobot(nullptr)
{
  construct(0.25, 0.14, aHDR);
}
PatternGeneratorInterface::~PatternGeneratorInterface()
{
  delete m_ZMPD;
  delete m_StepStackHandler;
  delete m_ZMPVRQP;
  delete m_PC;
  UnregisterPlugin(this);
}
PatternGeneratorInterface *patternGeneratorInterfaceFactory(CjrlHumanoidDynamicRobot *aHDR)
{
  return new PatternGeneratorInterface(aHDR);
}
void PatternGeneratorInterface::SetStartConfiguration(const COMState &com, const FootAbsolutePosition &lf,
                                                      const FootAbsolutePosition &rf)
{
  m_StartCOM = com; m_StartLF = lf; m_StartRF = rf;
}
int PatternGeneratorInterface::ParseCmd(std::istringstream &strm)
{
  std::string aCmd;
  strm >> aCmd;
  SimplePluginManager::CallMethod(aCmd, strm);
  return 0;
}
int PatternGeneratorInterface::initOnlineHerdt()
{
  // PGI.cpp:517-560 (the start configuration comes from SetStartConfiguration instead of the robot model)
  std::deque<RelativeFootPosition> rel;
  S3Vector zmp0;
  m_ZMPVRQP->SetCurrentTime(m_InternalClock);
  m_ZMPVRQP->InitOnLine(m_ZMPPositions, m_COMBuffer, m_LeftFootPositions, m_RightFootPositions, m_StartLF, m_StartRF, rel,
                        m_StartCOM, zmp0);
  m_Running = true;
  return 0;
}

// The first preview stage over the samples of the queues that have a full window behind them
// (ZMPPreviewControlWithMultiBodyZMP::FirstStageOfControl, ZMPPreviewControlWithMultiBodyZMP.cpp:378-446, for every tick at
// once): CoM x/y of sample k = OneIterationOfPreview on the window [k, k + NL).
void PatternGeneratorInterface::kajitaPreviewOverQueues()
{
  if (!m_PC->IsCoherent()) return;
  const size_t NL = (size_t)m_PC->Gains().NL, n = m_ZMPPositions.size();
  if (n < NL || m_PreviewedUpTo + NL > n) return;
  wg_ctx *ctx = default_context();
  walkgen_b200_check(wg_preview_set_gains(ctx, &m_PC->Gains()), "wg_preview_set_gains");
  const size_t first = m_PreviewedUpTo, cnt = n - first;          // samples [first, n): steps first .. n - NL
  std::vector<double> z(2 * cnt), com(6 * cnt, 0.0);
  for (size_t i = 0; i < cnt; ++i) { z[2 * i] = m_ZMPPositions[first + i].px; z[2 * i + 1] = m_ZMPPositions[first + i].py; }
  const int64_t off[2] = {0, (int64_t)cnt};
  wg_preview_plan *plan = nullptr;
  walkgen_b200_check(wg_preview_plan_create(ctx, 1, off, &plan), "wg_preview_plan_create");
  const int rc = wg_preview_run_batch(ctx, plan, WG_MEM_HOST, z.data(), m_PreviewState, com.data(), nullptr, 1);
  const int64_t steps = wg_preview_plan_total_steps(plan);
  wg_preview_plan_destroy(plan);
  walkgen_b200_check(rc, "wg_preview_run_batch");
  for (int64_t k = 0; k < steps; ++k) {
    COMState &c = m_COMBuffer[first + (size_t)k];
    for (int i = 0; i < 3; ++i) { c.x[i] = com[6 * k + i]; c.y[i] = com[6 * k + 3 + i]; }
  }
  m_PreviewedUpTo = first + (size_t)steps;
}

void PatternGeneratorInterface::ReadSequenceOfSteps(std::istringstream &strm)
{
  m_StepStackHandler->ReadStepSequenceAccordingToWalkMode(strm);   // PGI.cpp:1012-1028
}
void PatternGeneratorInterface::FinishAndRealizeStepSequence()
{
  // PGI.cpp:881-1005: CommonInitializationOfWalking (start configuration, copy + clear the step stack), CreateZMPReferences
  // (ZMPDiscretization::GetZMPDiscretization for the Kajita algorithms), strategy set-up, clock reset
  std::deque<RelativeFootPosition> rel;
  m_StepStackHandler->CopyRelativeFootPosition(rel, true);
  if (rel.empty()) return;
  m_ZMPD->SetZMPShift(m_ZMPShift);
  m_ZMPPositions.clear(); m_COMBuffer.clear(); m_LeftFootPositions.clear(); m_RightFootPositions.clear();
  S3Vector zmp0;
  m_ZMPD->SetComHeight(m_PC->GetHeightOfCoM() > 0.0 ? m_PC->GetHeightOfCoM() : m_StartCOM.z[0]);
  m_ZMPD->GetZMPDiscretization(m_ZMPPositions, m_COMBuffer, rel, m_LeftFootPositions, m_RightFootPositions, 0.0, m_StartCOM,
                               zmp0, m_StartLF, m_StartRF);
  for (double &v : m_PreviewState) v = 0.0;
  m_PreviewedUpTo = 0;
  kajitaPreviewOverQueues();
  m_KajitaOnLine = false;
  m_Running = true;
  m_InternalClock = 0.0;
}
void PatternGeneratorInterface::StartOnLineStepSequencing()
{
  // PGI.cpp:780-873: the stack must hold at least the first support foot and one step; the queues are started with
  // ZMPDiscretization::InitOnLine and refilled foot by foot from the 5 ms tick
  std::deque<RelativeFootPosition> rel;
  m_StepStackHandler->StartOnLineStep();
  if (m_StepStackHandler->ReturnStackSize() == 0) {   // default first steps of the reference's on-line mode
    m_StepStackHandler->PrepareForSupportFoot(-1);
    m_StepStackHandler->AddStandardOnLineStep(false, 0.0, 0.0, 0.0);
  }
  m_StepStackHandler->CopyRelativeFootPosition(rel, false);
  m_ZMPPositions.clear(); m_COMBuffer.clear(); m_LeftFootPositions.clear(); m_RightFootPositions.clear();
  S3Vector zmp0;
  m_ZMPD->SetZMPShift(m_ZMPShift);
  m_ZMPD->SetComHeight(m_PC->GetHeightOfCoM() > 0.0 ? m_PC->GetHeightOfCoM() : m_StartCOM.z[0]);
  m_ZMPD->InitOnLine(m_ZMPPositions, m_COMBuffer, m_LeftFootPositions, m_RightFootPositions, m_StartLF, m_StartRF, rel,
                     m_StartCOM, zmp0);
  for (double &v : m_PreviewState) v = 0.0;
  m_PreviewedUpTo = 0;
  kajitaPreviewOverQueues();
  m_KajitaOnLine = true;
  m_Running = true;
  m_InternalClock = 0.0;
}
void PatternGeneratorInterface::StopOnLineStepSequencing()
{
  // PGI.cpp:875-878: the stack is emptied and the transition flagged; the tick then takes one more default step (which
  // brings the feet side by side) and closes the walk with EndPhaseOfTheWalking (PGI.cpp:1283-1314)
  if (!m_KajitaOnLine) return;
  m_StepStackHandler->StopOnLineStep();
}
void PatternGeneratorInterface::AddOnLineStep(double X, double Y, double Theta)
{
  m_StepStackHandler->AddStandardOnLineStep(true, X, Y, Theta);
  if (!m_KajitaOnLine) return;
  RelativeFootPosition f = m_StepStackHandler->ReturnBackFootPosition();
  m_ZMPD->OnLineAddFoot(f, m_ZMPPositions, m_COMBuffer, m_LeftFootPositions, m_RightFootPositions, false);
  kajitaPreviewOverQueues();
}
void PatternGeneratorInterface::AddStepInStack(double dx, double dy, double theta)
{
  m_StepStackHandler->AddStepInTheStack(dx, dy, theta, m_StepStackHandler->GetSingleTimeSupport(),
                                        m_StepStackHandler->GetDoubleTimeSupport());
}

void PatternGeneratorInterface::CallMethod(std::string &Method, std::istringstream &strm)
{
  if (Method == ":SetAlgoForZmpTrajectory") {
    std::string algo;
    strm >> algo;
    m_AlgorithmforZMPCOM = (algo == "Herdt") ? 1 : 0;   // Kajita / KajitaOneStage / PBW / Morisawa / Dimitrov: the Kajita front end
  } else if (Method == ":HerdtOnline") {
    initOnlineHerdt();             // the handler takes no argument: the three numbers of the test are ignored (PGI.cpp:1103-1109)
  } else if (Method == ":setVelReference") {
    m_ZMPVRQP->Reference(strm);
  } else if (Method == ":setCoMPerturbationForce") {
    m_ZMPVRQP->setCoMPerturbationForce(strm);
  } else if (Method == ":stepseq") {
    ReadSequenceOfSteps(strm);                          // PGI.cpp:562-571 (m_StepSequence)
    FinishAndRealizeStepSequence();
  } else if (Method == ":finish") {
    FinishAndRealizeStepSequence();                     // PGI.cpp:1147
  } else if (Method == ":StartOnLineStepSequencing") {
    m_StepStackHandler->m_PartialStepSequence(strm);    // PGI.cpp:1150-1154
    StartOnLineStepSequencing();
  } else if (Method == ":StopOnLineStepSequencing") {
    StopOnLineStepSequencing();
  } else if (Method == ":ZMPShiftParameters") {
    for (int i = 0; i < 4 && !strm.eof(); ++i) strm >> m_ZMPShift[i];   // PGI.cpp:417-428
  } else if (Method == ":SetAutoFirstStep") {
    std::string v;
    strm >> v;
    m_AutoFirstStep = (v == "true");
  }
  // :LimitsFeasibility, :TimeDistributionParameters, :readfilefromkw, :ChangeNextStep configure subsystems outside the
  // accelerated path (step-over planner, KineoWorks files, Morisawa's on-line foot change): accepted, no effect
}

bool PatternGeneratorInterface::RunOneStepOfTheControlLoop(COMState &COMStateOut, ZMPPosition &ZMPTarget,
                                                           FootAbsolutePosition &LeftFootPosition,
                                                           FootAbsolutePosition &RightFootPosition)
{
  m_InternalClock += 0.005;        // PGI.cpp:1256
  if (!m_Running) return false;
  if (m_AlgorithmforZMPCOM == 1) {
    m_ZMPVRQP->OnLine(m_InternalClock, m_ZMPPositions, m_COMBuffer, m_LeftFootPositions, m_RightFootPositions);
    // CoMAndFootOnlyStrategy::OneGlobalStepOfControl, CoMAndFootOnlyStrategy.cpp:56-124
    if (m_ZMPPositions.empty() || m_COMBuffer.empty() || m_LeftFootPositions.empty() || m_RightFootPositions.empty()) {
      m_Running = false;
      return false;
    }
  } else {
    // DoubleStagePreviewControlStrategy: a tick needs two preview windows of ZMP reference ahead of it (first and second
    // stage, ZMPPreviewControlWithMultiBodyZMP.cpp:317-446): the motion ends when fewer than 2 NL samples are left
    // (TestKajita2003: 4002 samples discretised, 3362 ticks returned)
    const size_t NL = m_PC->IsCoherent() ? (size_t)m_PC->Gains().NL : (size_t)(1.6 / 0.005);
    if (m_KajitaOnLine && m_ZMPPositions.size() <= 2 * NL + 1) {
      // PGI.cpp:1283-1314: the queues run low in on-line mode: take the next foot from the stack (or a default step)
      RelativeFootPosition f;
      if (m_StepStackHandler->ReturnStackSize() == 0) m_StepStackHandler->AddStandardOnLineStep(false, 0.0, 0.0, 0.0);
      if (m_StepStackHandler->ReturnFrontFootPosition(f)) {
        const bool last = m_StepStackHandler->RemoveFirstStepInTheStack();   // true: stack empty after :StopOnLineStepSequencing
        m_ZMPD->OnLineAddFoot(f, m_ZMPPositions, m_COMBuffer, m_LeftFootPositions, m_RightFootPositions, last);
        if (last) m_KajitaOnLine = false;
        kajitaPreviewOverQueues();
      }
    }
    if (m_ZMPPositions.size() <= 2 * NL) {
      m_Running = false;
      return false;
    }
  }
  COMStateOut = m_COMBuffer.front(); ZMPTarget = m_ZMPPositions.front();
  LeftFootPosition = m_LeftFootPositions.front(); RightFootPosition = m_RightFootPositions.front();
  m_COMBuffer.pop_front(); m_ZMPPositions.pop_front(); m_LeftFootPositions.pop_front(); m_RightFootPositions.pop_front();
  if (m_AlgorithmforZMPCOM != 1 && m_PreviewedUpTo > 0) --m_PreviewedUpTo;
  return true;
}

// ---- the overloads of patterngeneratorinterface.hh:115-176.  Joint space is outside the accelerated path (no inverse
// kinematics, CoMAndFootOnlyStrategy semantics): configuration / velocity / acceleration are left as passed in.
bool PatternGeneratorInterface::RunOneStepOfTheControlLoop(MAL_VECTOR_TYPE(double) &CurrentConfiguration,
                                                           MAL_VECTOR_TYPE(double) &CurrentVelocity,
                                                           MAL_VECTOR_TYPE(double) &CurrentAcceleration,
                                                           MAL_VECTOR_TYPE(double) &ZMPTarget, COMState &finalCOMState,
                                                           FootAbsolutePosition &LeftFootPosition,
                                                           FootAbsolutePosition &RightFootPosition)
{
  (void)CurrentConfiguration; (void)CurrentVelocity; (void)CurrentAcceleration;
  ZMPPosition z;
  const bool r = RunOneStepOfTheControlLoop(finalCOMState, z, LeftFootPosition, RightFootPosition);
  if (r) {
    // PGI.cpp:1343-1353: the ZMP in the waist frame; the waist is the CoM on this path (CurrentConfiguration stays 0)
    ZMPTarget.resize(3);
    ZMPTarget[0] = z.px; ZMPTarget[1] = z.py; ZMPTarget[2] = z.pz;
  }
  return r;
}
bool PatternGeneratorInterface::RunOneStepOfTheControlLoop(MAL_VECTOR_TYPE(double) &CurrentConfiguration,
                                                           MAL_VECTOR_TYPE(double) &CurrentVelocity,
                                                           MAL_VECTOR_TYPE(double) &CurrentAcceleration,
                                                           MAL_VECTOR_TYPE(double) &ZMPTarget)
{
  COMState c; FootAbsolutePosition l, r;
  return RunOneStepOfTheControlLoop(CurrentConfiguration, CurrentVelocity, CurrentAcceleration, ZMPTarget, c, l, r);
}
bool PatternGeneratorInterface::RunOneStepOfTheControlLoop(FootAbsolutePosition &LeftFootPosition,
                                                           FootAbsolutePosition &RightFootPosition, ZMPPosition &ZMPRefPos,
                                                           COMPosition &COMRefPos)
{
  return RunOneStepOfTheControlLoop(COMRefPos, ZMPRefPos, LeftFootPosition, RightFootPosition);
}

}  // namespace PatternGeneratorJRL
