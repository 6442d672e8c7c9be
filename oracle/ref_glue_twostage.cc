/* oracle/ref_glue_twostage.cc - TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" handle onto the reference's own ZMPPreviewControlWithMultiBodyZMP object code (compiled by oracle/Makefile
 * from /root/reference/src/PreviewControl/ZMPPreviewControlWithMultiBodyZMP.cpp over the stand-in headers of
 * oracle/ref_shim).  The two collaborators that need jrl-dynamics and the HRP-2 model are replaced by test doubles that
 * hold no model: the "robot" returns the multibody ZMP of a caller-supplied stream (ref_shim/abstract-robot-dynamics),
 * and the ComAndFootRealization below only records which first-stage iteration was realised and the CoM it was given.
 * Everything between - the FIFOs, both preview stages, the delta ZMP, the NL-delayed sum - is the reference's code.
 */
#include <unistd.h>
#include <cstring>
#include <deque>
#include <sstream>
#include <string>
#include <vector>
using std::string;

#define private public
#define protected public
#include <PreviewControl/ZMPPreviewControlWithMultiBodyZMP.hh>
#undef private
#undef protected

using namespace PatternGeneratorJRL;

namespace {

class RecordingRealization : public ComAndFootRealization {
 public:
  RecordingRealization() : ComAndFootRealization(0), start_x(0), start_y(0), start_z(0) {}
  double start_x, start_y, start_z;
  std::vector<double> stage1;   /* [iteration][6] = (x, dx, ddx, y, dy, ddy) handed over at stage 0 */
  void Initialization() {}
  void CallMethod(std::string &, std::istringstream &) {}
  bool ComputePostureForGivenCoMAndFeetPosture(MAL_VECTOR_TYPE(double) &CoMPosition, MAL_VECTOR_TYPE(double) &CoMSpeed,
                                               MAL_VECTOR_TYPE(double) &CoMAcc, MAL_VECTOR_TYPE(double) &,
                                               MAL_VECTOR_TYPE(double) &, MAL_VECTOR_TYPE(double) &,
                                               MAL_VECTOR_TYPE(double) &, MAL_VECTOR_TYPE(double) &, int IterationNumber,
                                               int Stage)
  {
    if (Stage == 0) {
      getHumanoidDynamicRobot()->iteration = IterationNumber;
      if ((long)stage1.size() < 6 * ((long)IterationNumber + 1)) stage1.resize(6 * ((size_t)IterationNumber + 1), 0.0);
      double *o = &stage1[6 * (size_t)IterationNumber];
      o[0] = CoMPosition(0); o[1] = CoMSpeed(0); o[2] = CoMAcc(0);
      o[3] = CoMPosition(1); o[4] = CoMSpeed(1); o[5] = CoMAcc(1);
    }
    return true;
  }
  bool InitializationCoM(MAL_VECTOR_TYPE(double) &, MAL_S3_VECTOR_TYPE(double) &lStartingCOMPosition,
                         MAL_VECTOR_TYPE(double) &, FootAbsolutePosition &, FootAbsolutePosition &)
  {
    lStartingCOMPosition[0] = start_x; lStartingCOMPosition[1] = start_y; lStartingCOMPosition[2] = start_z;
    return true;
  }
  bool InitializationUpperBody(deque<ZMPPosition> &, deque<COMPosition> &, deque<RelativeFootPosition>) { return true; }
  MAL_S4x4_MATRIX_TYPE(double) GetCurrentPositionofWaistInCOMFrame() { return MAL_S4x4_MATRIX_TYPE(double)(); }
  MAL_S3_VECTOR_TYPE(double) GetCOGInitialAnkles() { return MAL_S3_VECTOR_TYPE(double)(); }
};

struct RefTwoStage {
  SimplePluginManager spm;
  CjrlHumanoidDynamicRobot robot;
  RecordingRealization cfr;
  ZMPPreviewControlWithMultiBodyZMP *zpc;
};

}  // namespace

namespace {
struct CwdGuard {     /* the constructor resets two debug files in the working directory (:72-74): keep them out of the repository */
  char old[4096];
  CwdGuard() { if (!getcwd(old, sizeof old)) old[0] = 0; if (chdir("/tmp")) {} }
  ~CwdGuard() { if (old[0] && chdir(old)) {} }
};
}  // namespace

extern "C" {

/* ctor (:46-95) creates its own PreviewControl (MODE_WITHOUT_INITIALPOS, auto weights); the three parameters are set
 * through the PreviewControl setters and SetPreviewControl (:101-108) re-reads them (m_NL). */
void *ref_twostage_new(double T, double preview_time, double zc)
{
  RefTwoStage *h = new RefTwoStage;
  { CwdGuard g; h->zpc = new ZMPPreviewControlWithMultiBodyZMP(&h->spm); }
  PreviewControl *pc = h->zpc->m_PC;
  pc->SetSamplingPeriod(T);
  pc->SetPreviewControlTime(preview_time);
  pc->SetHeightOfCoM(zc);
  h->zpc->SetPreviewControl(pc);
  h->cfr.setHumanoidDynamicRobot(&h->robot);
  h->zpc->setComAndFootRealization(&h->cfr);
  h->zpc->setHumanoidDynamicRobot(&h->robot);
  return h;
}
void ref_twostage_delete(void *hv)
{
  RefTwoStage *h = static_cast<RefTwoStage *>(hv);
  delete h->zpc;
  delete h;
}
/* the PreviewControl the object owns (for ref_preview_get_gains of ref_glue_preview.cc, which takes a RefPreview handle:
 * this returns the bare PreviewControl *, see ref_twostage_get_gains) */
int ref_twostage_get_gains(void *hv, double *A9, double *B3, double *C3, double *Kx3, double *Ks, double *F, int capF)
{
  PreviewControl *pc = static_cast<RefTwoStage *>(hv)->zpc->m_PC;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) A9[3 * i + j] = pc->m_A(i, j);
    B3[i] = pc->m_B(i, 0);
    C3[i] = pc->m_C(0, i);
    Kx3[i] = pc->m_Kx(0, i);
  }
  *Ks = pc->m_Ks;
  int nl = (int)pc->m_SizeOfPreviewWindow;
  for (int i = 0; i < nl && i < capF && i < (int)pc->m_F.size1(); ++i) F[i] = pc->m_F(i, 0);
  return nl;
}

/* EvaluateStartingCoM (:812-826), Setup (:487-528: ComputeOptimalWeights with the reference's own dgges_ path, then
 * SetupFirstPhase + NL x SetupIterativePhase), then OneGlobalStepOfControl (:194-266) + UpdateTheZMPRefQueue (:753) until
 * the reference stream or the multibody stream runs out.  strategy: ZMPCOM_TRAJECTORY_FULL = 1, SECOND_STAGE_ONLY = 2,
 * FIRST_STAGE_ONLY = 3 (:SetStrategyForStageActivation).  Outputs as oracle_two_stage_run; returns the number of
 * OneGlobalStepOfControl calls, -1 when the reference throws. */
long ref_twostage_run(void *hv, const double *zmpref_xy, long L, const double *zmb, long n_zmb, const double *com_start_xyz,
                      int strategy, double *stage1, double *delta, double *final_com, long *ticks_out)
{
  RefTwoStage *h = static_cast<RefTwoStage *>(hv);
  ZMPPreviewControlWithMultiBodyZMP *z = h->zpc;
  const long NL = (long)z->m_NL;
  if (L < 2 * NL + 1) return -1;
  h->robot.zmp_stream = zmb; h->robot.zmp_stream_len = n_zmb; h->robot.iteration = 0;
  h->cfr.stage1.clear();
  h->cfr.start_x = com_start_xyz[0]; h->cfr.start_y = com_start_xyz[1]; h->cfr.start_z = com_start_xyz[2];
  z->SetStrategyForStageActivation(strategy);
  std::deque<ZMPPosition> ref(L);
  std::deque<COMState> coms(L);
  std::deque<FootAbsolutePosition> lf(L), rf(L);
  for (long i = 0; i < L; ++i) {
    std::memset(&ref[i], 0, sizeof(ZMPPosition));
    std::memset(&coms[i], 0, sizeof(COMState));
    std::memset(&lf[i], 0, sizeof(FootAbsolutePosition));
    std::memset(&rf[i], 0, sizeof(FootAbsolutePosition));
    ref[i].px = zmpref_xy[2 * i]; ref[i].py = zmpref_xy[2 * i + 1]; ref[i].time = 0.005 * i;
    coms[i].z[0] = com_start_xyz[2];
  }
  long steps = 0, tick = 0;
  try {
    MAL_VECTOR_DIM(body, double, 36);
    MAL_S3_VECTOR(sc, double);
    MAL_VECTOR_DIM(waist, double, 6);
    z->EvaluateStartingCoM(body, sc, waist, lf[0], rf[0]);
    if (n_zmb < NL) return -1;
    z->Setup(ref, coms, lf, rf);
    tick = NL;
    if (delta)
      for (long i = 0; i < NL; ++i) { delta[2 * i] = z->m_FIFODeltaZMPPositions[i].px; delta[2 * i + 1] = z->m_FIFODeltaZMPPositions[i].py; }
    MAL_VECTOR_DIM(q, double, 36);
    MAL_VECTOR_DIM(dq, double, 36);
    MAL_VECTOR_DIM(ddq, double, 36);
    long next_ref = 2 * NL + 1;
    while ((long)z->m_FIFOZMPRefPositions.size() >= NL && tick < n_zmb) {
      COMState fin; std::memset(&fin, 0, sizeof fin);
      fin.z[0] = com_start_xyz[2];
      ZMPPosition zp; std::memset(&zp, 0, sizeof zp);
      z->OneGlobalStepOfControl(lf[0], rf[0], zp, fin, q, dq, ddq);
      if (delta && strategy != 3) {
        const ZMPPosition &d = z->m_FIFODeltaZMPPositions.back();
        delta[2 * tick] = d.px; delta[2 * tick + 1] = d.py;
      }
      if (final_com) {
        double *o = final_com + 6 * steps;
        o[0] = fin.x[0]; o[1] = fin.x[1]; o[2] = fin.x[2]; o[3] = fin.y[0]; o[4] = fin.y[1]; o[5] = fin.y[2];
      }
      ++steps; ++tick;
      if (next_ref < L) { z->UpdateTheZMPRefQueue(ref[next_ref]); ++next_ref; }
    }
  } catch (...) {
    return -1;
  }
  if (stage1) {
    const long n = (long)h->cfr.stage1.size() / 6;
    for (long i = 0; i < 6 * (n < tick ? n : tick); ++i) stage1[i] = h->cfr.stage1[i];
  }
  if (ticks_out) *ticks_out = tick;
  return steps;
}

}  // extern "C"
