// walkgen_host_twostage.cpp - class mirror of ZMPPreviewControlWithMultiBodyZMP (Kajita's two-stage scheme) over the C ABI.
// FIFO bookkeeping follows src/PreviewControl/ZMPPreviewControlWithMultiBodyZMP.cpp statement for statement; the two preview
// iterations of a tick run on the GPU through the PreviewControl mirror, the batched form through wg_preview_run_batch /
// wg_preview_delta_zmp / wg_preview_stage2_run_batch.  The posture realisation and the multibody ZMP are the caller's.
#include "walkgen_host.hh"
#include <sstream>

namespace PatternGeneratorJRL {

static void check2(int rc, const char *what)
{
  if (rc != WG_OK) throw std::runtime_error(std::string(what) + ": " + wg_last_error(walkgen_b200::default_context()));
}

ZMPPreviewControlWithMultiBodyZMP::ZMPPreviewControlWithMultiBodyZMP(SimplePluginManager *lSPM)
    : SimplePlugin(lSPM), m_PC(0), m_OwnPC(true), m_HumanoidDynamicRobot(0), m_ComAndFootRealization(0), m_sxzmp(0),
      m_syzmp(0), m_sxDeltazmp(0), m_syDeltazmp(0), m_SamplingPeriod(-1), m_PreviewControlTime(0), m_NL(0),
      m_StageStrategy(ZMPCOM_TRAJECTORY_FULL), m_NumberOfIterations(0), m_StartingNewSequence(true)
{
  // ctor, :46-95: registers its three methods, sizes the four 3 x 1 states, owns a PreviewControl(WITHOUT_INITIALPOS, auto)
  std::string names[3] = {":samplingperiod", ":previewcontroltime", ":comheight"};
  for (int i = 0; i < 3; ++i) RegisterMethod(names[i]);
  m_PC1x.resize(3, 1); m_PC1y.resize(3, 1); m_Deltax.resize(3, 1); m_Deltay.resize(3, 1);
  m_PC = new PreviewControl(lSPM, OptimalControllerSolver::MODE_WITHOUT_INITIALPOS, true);
}

ZMPPreviewControlWithMultiBodyZMP::~ZMPPreviewControlWithMultiBodyZMP()
{
  if (m_OwnPC) delete m_PC;
}

void ZMPPreviewControlWithMultiBodyZMP::SetPreviewControl(PreviewControl *aPC)   // :101-108
{
  if (m_OwnPC && m_PC != aPC) { delete m_PC; m_OwnPC = false; }
  m_PC = aPC;
  m_SamplingPeriod = m_PC->SamplingPeriod();
  m_PreviewControlTime = m_PC->PreviewControlTime();
  m_NL = (unsigned int)(m_PreviewControlTime / m_SamplingPeriod);
}

void ZMPPreviewControlWithMultiBodyZMP::SetStrategyForStageActivation(int aZMPComTraj)   // :758-776
{
  switch (aZMPComTraj) {
  case ZMPCOM_TRAJECTORY_FULL: m_StageStrategy = ZMPCOM_TRAJECTORY_FULL; break;
  case ZMPCOM_TRAJECTORY_SECOND_STAGE_ONLY: m_StageStrategy = ZMPCOM_TRAJECTORY_SECOND_STAGE_ONLY; break;
  case ZMPCOM_TRAJECTORY_FIRST_STAGE_ONLY: m_StageStrategy = ZMPCOM_TRAJECTORY_FIRST_STAGE_ONLY; break;
  default: break;
  }
}

void ZMPPreviewControlWithMultiBodyZMP::CallToComAndFootRealization(COMState &acomp, FootAbsolutePosition &aLeftFAP,
                                                                    FootAbsolutePosition &aRightFAP,
                                                                    MAL_VECTOR_TYPE(double) &CurrentConfiguration,
                                                                    MAL_VECTOR_TYPE(double) &CurrentVelocity,
                                                                    MAL_VECTOR_TYPE(double) &CurrentAcceleration,
                                                                    int IterationNumber, int StageOfTheAlgorithm)   // :110-192
{
  std::vector<double> aCOMState(6), aCOMSpeed(6), aCOMAcc(6), aLeftFootPosition(5), aRightFootPosition(5);
  aCOMState[0] = acomp.x[0]; aCOMState[1] = acomp.y[0]; aCOMState[2] = acomp.z[0];
  aCOMState[3] = acomp.roll[0]; aCOMState[4] = acomp.pitch[0]; aCOMState[5] = acomp.yaw[0];
  aCOMSpeed[0] = acomp.x[1]; aCOMSpeed[1] = acomp.y[1]; aCOMSpeed[2] = acomp.z[1];
  aCOMSpeed[3] = acomp.roll[1]; aCOMSpeed[4] = acomp.roll[1]; aCOMSpeed[5] = acomp.roll[1];   // sic, :135-137
  aCOMAcc[0] = acomp.x[2]; aCOMAcc[1] = acomp.y[2]; aCOMAcc[2] = acomp.z[2];
  aCOMAcc[3] = acomp.roll[2]; aCOMAcc[4] = acomp.roll[2]; aCOMAcc[5] = acomp.roll[2];         // sic, :142-144
  aLeftFootPosition[0] = aLeftFAP.x; aLeftFootPosition[1] = aLeftFAP.y; aLeftFootPosition[2] = aLeftFAP.z;
  aLeftFootPosition[3] = aLeftFAP.theta; aLeftFootPosition[4] = aLeftFAP.omega;
  aRightFootPosition[0] = aRightFAP.x; aRightFootPosition[1] = aRightFAP.y; aRightFootPosition[2] = aRightFAP.z;
  aRightFootPosition[3] = aRightFAP.theta; aRightFootPosition[4] = aRightFAP.omega;
  if (m_HumanoidDynamicRobot) {
    CurrentConfiguration = m_HumanoidDynamicRobot->currentConfiguration();
    CurrentVelocity = m_HumanoidDynamicRobot->currentVelocity();
    CurrentAcceleration = m_HumanoidDynamicRobot->currentAcceleration();
  }
  if (m_ComAndFootRealization)
    m_ComAndFootRealization->ComputePostureForGivenCoMAndFeetPosture(aCOMState, aCOMSpeed, aCOMAcc, aLeftFootPosition,
                                                                     aRightFootPosition, CurrentConfiguration,
                                                                     CurrentVelocity, CurrentAcceleration, IterationNumber,
                                                                     StageOfTheAlgorithm);
  if (StageOfTheAlgorithm == 0 && m_HumanoidDynamicRobot) {
    m_HumanoidDynamicRobot->currentConfiguration(CurrentConfiguration);
    m_HumanoidDynamicRobot->currentVelocity(CurrentVelocity);
    m_HumanoidDynamicRobot->currentAcceleration(CurrentAcceleration);
  }
}

int ZMPPreviewControlWithMultiBodyZMP::OneGlobalStepOfControl(FootAbsolutePosition &LeftFootPosition,
                                                              FootAbsolutePosition &RightFootPosition, ZMPPosition &,
                                                              COMState &refandfinalCOMState,
                                                              MAL_VECTOR_TYPE(double) &CurrentConfiguration,
                                                              MAL_VECTOR_TYPE(double) &CurrentVelocity,
                                                              MAL_VECTOR_TYPE(double) &CurrentAcceleration)   // :194-307
{
  FirstStageOfControl(LeftFootPosition, RightFootPosition, refandfinalCOMState);
  COMState acompos = m_FIFOCOMStates[m_NL];
  FootAbsolutePosition aLeftFAP = m_FIFOLeftFootPosition[m_NL];
  FootAbsolutePosition aRightFAP = m_FIFORightFootPosition[m_NL];
  CallToComAndFootRealization(acompos, aLeftFAP, aRightFAP, CurrentConfiguration, CurrentVelocity, CurrentAcceleration,
                              m_NumberOfIterations, 0);
  if (m_StageStrategy != ZMPCOM_TRAJECTORY_FIRST_STAGE_ONLY) EvaluateMultiBodyZMP(-1);
  aLeftFAP = m_FIFOLeftFootPosition[0];
  aRightFAP = m_FIFORightFootPosition[0];
  SecondStageOfControl(refandfinalCOMState);
  if (m_StageStrategy != ZMPCOM_TRAJECTORY_FIRST_STAGE_ONLY)
    CallToComAndFootRealization(refandfinalCOMState, aLeftFAP, aRightFAP, CurrentConfiguration, CurrentVelocity,
                                CurrentAcceleration, m_NumberOfIterations - (int)m_NL, 1);
  m_NumberOfIterations++;
  return 1;
}

int ZMPPreviewControlWithMultiBodyZMP::SecondStageOfControl(COMState &finalCOMState)   // :317-376
{
  double Deltazmpx2, Deltazmpy2;
  COMState aCOMState = m_FIFOCOMStates[0];
  const bool second = m_StageStrategy == ZMPCOM_TRAJECTORY_SECOND_STAGE_ONLY || m_StageStrategy == ZMPCOM_TRAJECTORY_FULL;
  if (second) {
    m_PC->OneIterationOfPreview(m_Deltax, m_Deltay, m_sxDeltazmp, m_syDeltazmp, m_FIFODeltaZMPPositions, 0, Deltazmpx2,
                                Deltazmpy2, true);
    // the CoM of NL ticks ago gets the correction, on all three derivatives
    for (int i = 0; i < 3; ++i) { aCOMState.x[i] += m_Deltax(i, 0); aCOMState.y[i] += m_Deltay(i, 0); }
  }
  finalCOMState = aCOMState;
  if (second) m_FIFODeltaZMPPositions.pop_front();
  m_FIFOCOMStates.pop_front();
  m_FIFOLeftFootPosition.pop_front();
  m_FIFORightFootPosition.pop_front();
  return 1;
}

int ZMPPreviewControlWithMultiBodyZMP::FirstStageOfControl(FootAbsolutePosition &LeftFootPosition,
                                                           FootAbsolutePosition &RightFootPosition, COMState &afCOMState)   // :378-446
{
  double zmpx2, zmpy2;
  COMState acomp;
  acomp.yaw[0] = 0.0; acomp.pitch[0] = 0.0;
  if (m_StageStrategy == ZMPCOM_TRAJECTORY_FULL || m_StageStrategy == ZMPCOM_TRAJECTORY_FIRST_STAGE_ONLY) {
    m_PC->OneIterationOfPreview(m_PC1x, m_PC1y, m_sxzmp, m_syzmp, m_FIFOZMPRefPositions, 0, zmpx2, zmpy2, true);
    for (unsigned j = 0; j < 3; ++j) {
      acomp.x[j] = m_PC1x(j, 0); acomp.y[j] = m_PC1y(j, 0); acomp.z[j] = afCOMState.z[j];
      acomp.yaw[j] = afCOMState.yaw[j]; acomp.pitch[j] = afCOMState.pitch[j]; acomp.roll[j] = afCOMState.roll[j];
    }
  } else if (m_StageStrategy == ZMPCOM_TRAJECTORY_SECOND_STAGE_ONLY) {
    for (unsigned j = 0; j < 3; ++j) {
      acomp.x[j] = m_PC1x(j, 0) = afCOMState.x[j]; acomp.y[j] = m_PC1y(j, 0) = afCOMState.y[j];
      acomp.z[j] = afCOMState.z[j]; acomp.yaw[j] = afCOMState.yaw[j]; acomp.pitch[j] = afCOMState.pitch[j];
    }
  }
  m_FIFOCOMStates.push_back(acomp);
  m_FIFORightFootPosition.push_back(RightFootPosition);
  m_FIFOLeftFootPosition.push_back(LeftFootPosition);
  m_FIFOZMPRefPositions.pop_front();
  return 1;
}

int ZMPPreviewControlWithMultiBodyZMP::EvaluateMultiBodyZMP(int)   // :447-485
{
  vector3d ZMPmultibody;
  ZMPmultibody[0] = ZMPmultibody[1] = ZMPmultibody[2] = 0.0;
  if (m_HumanoidDynamicRobot) {
    std::string sComputeZMP("ComputeBackwardDynamics"), sZMPtrue("true");
    m_HumanoidDynamicRobot->setProperty(sComputeZMP, sZMPtrue);
    m_HumanoidDynamicRobot->computeForwardKinematics();
    ZMPmultibody = m_HumanoidDynamicRobot->zeroMomentumPoint();
  }
  ZMPPosition aZMPpos;
  aZMPpos.px = m_FIFOZMPRefPositions[0].px - ZMPmultibody[0];
  aZMPpos.py = m_FIFOZMPRefPositions[0].py - ZMPmultibody[1];
  aZMPpos.pz = 0.0; aZMPpos.theta = 0.0; aZMPpos.stepType = 1;
  aZMPpos.time = m_FIFOZMPRefPositions[0].time;
  m_FIFODeltaZMPPositions.push_back(aZMPpos);
  m_StartingNewSequence = false;
  return 1;
}

int ZMPPreviewControlWithMultiBodyZMP::Setup(std::deque<ZMPPosition> &ZMPRefPositions, std::deque<COMState> &COMStates,
                                             std::deque<FootAbsolutePosition> &LeftFootPositions,
                                             std::deque<FootAbsolutePosition> &RightFootPositions)   // :487-528
{
  m_NumberOfIterations = 0;
  std::vector<double> q, dq, ddq;
  if (m_HumanoidDynamicRobot) {
    q = m_HumanoidDynamicRobot->currentConfiguration(); dq = m_HumanoidDynamicRobot->currentVelocity();
    ddq = m_HumanoidDynamicRobot->currentAcceleration();
  }
  m_PC->ComputeOptimalWeights(OptimalControllerSolver::MODE_WITHOUT_INITIALPOS);
  if (m_HumanoidDynamicRobot) {
    std::string inProperty[5] = {"TimeStep", "ComputeAcceleration", "ComputeBackwardDynamics", "ComputeZMP", "ResetIteration"};
    std::ostringstream oss; oss << m_SamplingPeriod;
    std::string inValue[5] = {oss.str(), "false", "false", "true", "true"};
    for (unsigned i = 0; i < 5; ++i) m_HumanoidDynamicRobot->setProperty(inProperty[i], inValue[i]);
  }
  SetupFirstPhase(ZMPRefPositions, COMStates, LeftFootPositions, RightFootPositions);
  for (unsigned i = 0; i < m_NL; ++i)
    SetupIterativePhase(ZMPRefPositions, COMStates, LeftFootPositions, RightFootPositions, q, dq, ddq, (int)i);
  return 0;
}

int ZMPPreviewControlWithMultiBodyZMP::SetupFirstPhase(std::deque<ZMPPosition> &ZMPRefPositions, std::deque<COMState> &,
                                                       std::deque<FootAbsolutePosition> &LeftFootPositions,
                                                       std::deque<FootAbsolutePosition> &RightFootPositions)   // :530-600
{
  if (ZMPRefPositions.size() < m_NL || LeftFootPositions.size() < m_NL || RightFootPositions.size() < m_NL)
    throw std::runtime_error("ZMPPreviewControlWithMultiBodyZMP::SetupFirstPhase: fewer samples than the preview window");
  m_sxzmp = 0.0; m_syzmp = 0.0; m_sxDeltazmp = 0.0; m_syDeltazmp = 0.0;
  m_StartingNewSequence = true;
  m_FIFOZMPRefPositions.resize(m_NL); m_FIFOLeftFootPosition.resize(m_NL); m_FIFORightFootPosition.resize(m_NL);
  for (unsigned i = 0; i < m_NL; ++i) {
    m_FIFOZMPRefPositions[i] = ZMPRefPositions[i];
    m_FIFOLeftFootPosition[i] = LeftFootPositions[i];
    m_FIFORightFootPosition[i] = RightFootPositions[i];
  }
  m_PC1x(0, 0) = m_StartingCOMState[0]; m_PC1x(1, 0) = 0.0; m_PC1x(2, 0) = 0.0;
  m_PC1y(0, 0) = m_StartingCOMState[1]; m_PC1y(1, 0) = 0.0; m_PC1y(2, 0) = 0.0;
  for (int i = 0; i < 3; ++i) { m_Deltax(i, 0) = 0.0; m_Deltay(i, 0) = 0.0; }
  m_FIFODeltaZMPPositions.clear();
  m_FIFOCOMStates.clear();
  if (m_HumanoidDynamicRobot) {
    std::vector<double> dq = m_HumanoidDynamicRobot->currentVelocity(), ddq = m_HumanoidDynamicRobot->currentAcceleration();
    for (size_t i = 0; i < dq.size(); ++i) dq[i] = 0.0;
    for (size_t i = 0; i < ddq.size(); ++i) ddq[i] = 0.0;
    m_HumanoidDynamicRobot->currentVelocity(dq);
    m_HumanoidDynamicRobot->currentAcceleration(ddq);
  }
  return 0;
}

int ZMPPreviewControlWithMultiBodyZMP::SetupIterativePhase(std::deque<ZMPPosition> &ZMPRefPositions,
                                                           std::deque<COMState> &COMStates,
                                                           std::deque<FootAbsolutePosition> &LeftFootPositions,
                                                           std::deque<FootAbsolutePosition> &RightFootPositions,
                                                           MAL_VECTOR_TYPE(double) &CurrentConfiguration,
                                                           MAL_VECTOR_TYPE(double) &CurrentVelocity,
                                                           MAL_VECTOR_TYPE(double) &CurrentAcceleration, int localindex)   // :602-664
{
  FirstStageOfControl(LeftFootPositions[localindex], RightFootPositions[localindex], COMStates[localindex]);
  // sic: right and left feet swapped in this call of the reference (:640-642)
  CallToComAndFootRealization(m_FIFOCOMStates[localindex], m_FIFORightFootPosition[localindex],
                              m_FIFOLeftFootPosition[localindex], CurrentConfiguration, CurrentVelocity,
                              CurrentAcceleration, m_NumberOfIterations, 0);
  EvaluateMultiBodyZMP(localindex);
  m_FIFOZMPRefPositions.push_back(ZMPRefPositions[localindex + 1 + m_NL]);   // ZMPRefPositions[NL] never enters the FIFO
  m_NumberOfIterations++;
  return 0;
}

void ZMPPreviewControlWithMultiBodyZMP::CreateExtraCOMBuffer(std::deque<COMState> &ExtraCOMBuffer,
                                                             std::deque<ZMPPosition> &ExtraZMPBuffer,
                                                             std::deque<ZMPPosition> &ExtraZMPRefBuffer)   // :666-751
{
  // a copy of the first stage run ahead over the extra reference: FIFO = Extra[0, NL), then per sample i push Extra[i],
  // one preview iteration at lindex 0, pop.  The stream is known in advance, so it is ONE batched pass.
  std::deque<ZMPPosition> stream;
  for (unsigned i = 0; i < m_NL; ++i) stream.push_back(ExtraZMPRefBuffer[i]);
  for (size_t i = 0; i < ExtraCOMBuffer.size(); ++i) stream.push_back(ExtraZMPRefBuffer[i]);
  walkgen_b200::Matrix aPC1x = m_PC1x, aPC1y = m_PC1y;
  double aSxzmp = m_sxzmp, aSyzmp = m_syzmp;
  std::vector<double> com6, zmp2;
  // window of sample i = stream[i, i + NL) needs NL + 1 entries at the time of the call (push before the iteration)
  m_PC->RunWholeTrajectory(stream, aPC1x, aPC1y, aSxzmp, aSyzmp, com6, zmp2, true);
  for (size_t i = 0; i < ExtraCOMBuffer.size(); ++i) {
    for (unsigned j = 0; j < 3; ++j) { ExtraCOMBuffer[i].x[j] = com6[6 * i + j]; ExtraCOMBuffer[i].y[j] = com6[6 * i + 3 + j]; }
    ExtraZMPBuffer[i].px = zmp2[2 * i]; ExtraZMPBuffer[i].py = zmp2[2 * i + 1];
    ExtraCOMBuffer[i].yaw[0] = ExtraZMPRefBuffer[i].theta;
  }
}

int ZMPPreviewControlWithMultiBodyZMP::EvaluateStartingCoM(MAL_VECTOR_TYPE(double) &BodyAnglesInit,
                                                           MAL_S3_VECTOR_TYPE(double) &aStartingCOMState,
                                                           MAL_VECTOR_TYPE(double) &aStartingWaistPosition,
                                                           FootAbsolutePosition &InitLeftFootPosition,
                                                           FootAbsolutePosition &InitRightFootPosition)   // :812-826
{
  if (m_ComAndFootRealization)
    m_ComAndFootRealization->InitializationCoM(BodyAnglesInit, m_StartingCOMState, aStartingWaistPosition,
                                               InitLeftFootPosition, InitRightFootPosition);
  aStartingCOMState[0] = m_StartingCOMState[0];
  aStartingCOMState[1] = m_StartingCOMState[1];
  aStartingCOMState[2] = m_StartingCOMState[2];
  return 0;
}

int ZMPPreviewControlWithMultiBodyZMP::EvaluateStartingState(MAL_VECTOR_TYPE(double) &BodyAnglesInit,
                                                             MAL_S3_VECTOR_TYPE(double) &aStartingCOMState,
                                                             MAL_S3_VECTOR_TYPE(double) &aStartingZMPPosition,
                                                             MAL_VECTOR_TYPE(double) &aStartingWaistPosition,
                                                             FootAbsolutePosition &InitLeftFootPosition,
                                                             FootAbsolutePosition &InitRightFootPosition)   // :797-810
{
  const int r = EvaluateStartingCoM(BodyAnglesInit, aStartingCOMState, aStartingWaistPosition, InitLeftFootPosition,
                                    InitRightFootPosition);
  if (m_ComAndFootRealization) aStartingZMPPosition = m_ComAndFootRealization->GetCOGInitialAnkles();
  return r;
}

void ZMPPreviewControlWithMultiBodyZMP::CallMethod(std::string &Method, std::istringstream &strm)   // :868-890
{
  if (Method == ":samplingperiod") {
    std::string a; strm >> a; m_SamplingPeriod = atof(a.c_str());
  } else if (Method == ":previewcontroltime") {
    std::string a; strm >> a; m_PreviewControlTime = atof(a.c_str());
  }
  if (m_SamplingPeriod > 0.0) m_NL = (unsigned int)(m_PreviewControlTime / m_SamplingPeriod);   // :864-866 set m_NL = 0 first
}

int ZMPPreviewControlWithMultiBodyZMP::RunWholeTrajectory(std::deque<ZMPPosition> &Z, const COMState &StartingCOM,
                                                          void (*multibody_zmp)(void *, long, const double *, double *),
                                                          void *user, std::deque<COMState> &Final)
{
  const size_t NL = m_NL, L = Z.size();
  Final.clear();
  if (L < 2 * NL + 1 || !multibody_zmp) return 0;
  m_PC->ComputeOptimalWeights(OptimalControllerSolver::MODE_WITHOUT_INITIALPOS);   // Setup, :497
  m_PC->BindGains();
  wg_ctx *ctx = walkgen_b200::default_context();
  // the stream the FIFO holds: ZMPRefPositions[NL] skipped (:660)
  const size_t Le = L - 1;
  std::vector<double> zeff(2 * Le), com1(6 * Le, 0.0), zmb(2 * Le, 0.0), delta(2 * Le, 0.0), fin(6 * Le, 0.0);
  for (size_t k = 0; k < Le; ++k) { const ZMPPosition &p = Z[k < NL ? k : k + 1]; zeff[2 * k] = p.px; zeff[2 * k + 1] = p.py; }
  const int64_t offs[2] = {0, (int64_t)Le};
  wg_preview_plan *plan = nullptr;
  check2(wg_preview_plan_create(ctx, 1, offs, &plan), "wg_preview_plan_create");
  double st1[8] = {StartingCOM.x[0], 0, 0, StartingCOM.y[0], 0, 0, 0, 0}, st2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int rc = wg_preview_run_batch(ctx, plan, WG_MEM_HOST, zeff.data(), st1, com1.data(), nullptr, 1);
  const size_t ticks = Le - NL + 1;
  if (rc == WG_OK) {
    for (size_t k = 0; k < ticks; ++k) multibody_zmp(user, (long)k, &com1[6 * k], &zmb[2 * k]);
    rc = wg_preview_delta_zmp(ctx, plan, WG_MEM_HOST, zeff.data(), zmb.data(), delta.data());
  }
  if (rc == WG_OK) {
    if (m_StageStrategy == ZMPCOM_TRAJECTORY_FIRST_STAGE_ONLY) fin = com1;
    else rc = wg_preview_stage2_run_batch(ctx, plan, WG_MEM_HOST, delta.data(), com1.data(), st2, fin.data(), nullptr);
  }
  wg_preview_plan_destroy(plan);
  check2(rc, "ZMPPreviewControlWithMultiBodyZMP::RunWholeTrajectory");
  const size_t steps = Le - 2 * NL + 1;
  for (size_t n = 0; n < steps; ++n) {
    COMState c = StartingCOM;
    for (int j = 0; j < 3; ++j) { c.x[j] = fin[6 * n + j]; c.y[j] = fin[6 * n + 3 + j]; }
    Final.push_back(c);
  }
  return (int)steps;
}

}  // namespace PatternGeneratorJRL
