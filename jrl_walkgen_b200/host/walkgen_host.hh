// walkgen_host.hh - host-side C++ mirror of the reference's class interfaces for the accelerated hot path, written
// over the C ABI of include/walkgen_b200.h (batch size 1 unless stated).  Same class names, method names, argument
// meaning and command strings as the reference, so that callers written against jrl-walkgen (its PGI, its tests)
// read the same:
//   SimplePlugin / SimplePluginManager        src/SimplePlugin.hh:46-72, src/SimplePluginManager.{hh,cpp}
//   PreviewControl                            src/PreviewControl/PreviewControl.hh:58-140
//   OptCholesky                               src/Mathematics/OptCholesky.hh:55-102
//   Optimization::Solver::PLDPSolver          src/Mathematics/PLDPSolver.hh:48-68
//   ZMPRefTrajectoryGeneration                src/ZMPRefTrajectoryGeneration/ZMPRefTrajectoryGeneration.hh:208-328
//   ZMPVelocityReferencedQP                   src/ZMPRefTrajectoryGeneration/ZMPVelocityReferencedQP.hh:59-131
//   PatternGeneratorInterface (Herdt path)    include/jrl/walkgen/patterngeneratorinterface.hh:55-306
// The jrl-mal matrix type is replaced by a tiny dense matrix with the same element access (MAL_MATRIX macros below).
// There is no CPU fallback: every class needs a CUDA device and throws std::runtime_error without one.
#ifndef WALKGEN_B200_HOST_HH
#define WALKGEN_B200_HOST_HH

#include <deque>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <cstring>
#include "../../include/walkgen_b200.h"

namespace walkgen_b200 {

/* Process-wide context on device $WG_DEVICE (default 0), created on first use. */
wg_ctx *default_context();

/* Minimal stand-in for MAL_MATRIX(name,double): row-major dense matrix with (i,j) access. */
class Matrix {
 public:
  Matrix() : r_(0), c_(0) {}
  Matrix(unsigned r, unsigned c) : r_(r), c_(c), d_(r * c, 0.0) {}
  void resize(unsigned r, unsigned c) { r_ = r; c_ = c; d_.assign((size_t)r * c, 0.0); }
  double &operator()(unsigned i, unsigned j) { return d_[(size_t)i * c_ + j]; }
  double operator()(unsigned i, unsigned j) const { return d_[(size_t)i * c_ + j]; }
  unsigned size1() const { return r_; }
  unsigned size2() const { return c_; }
  double *data() { return d_.data(); }
 private:
  unsigned r_, c_;
  std::vector<double> d_;
};

}  // namespace walkgen_b200

/* ---- stand-ins for the abstract-robot-dynamics interfaces the hot path reads ------------------------------------------
 * The reference asks its CjrlHumanoidDynamicRobot for exactly this on the accelerated path: the sole size and ankle
 * position of the feet (relative-feet-inequalities.cpp:153-182, FootConstraintsAsLinearSystem.cpp:269-281) and the bounds
 * of the hip-yaw joints, found as jointsBetween(waist, ankle)[1] (OrientationsPreview.cpp:42-68).  A build that has the real
 * abstract-robot-dynamics headers defines WALKGEN_B200_HAVE_ABSTRACT_ROBOT_DYNAMICS and passes its own robot. */
#ifndef WALKGEN_B200_HAVE_ABSTRACT_ROBOT_DYNAMICS
struct vector3d {
  double v[3];
  double &operator[](int i) { return v[i]; }
  const double &operator[](int i) const { return v[i]; }
};
class CjrlJoint {
 public:
  CjrlJoint(double lower = 0.0, double upper = 0.0, double upperVelocity = 0.0) : m_Lo(lower), m_Up(upper), m_Vel(upperVelocity) {}
  double lowerBound(unsigned int) const { return m_Lo; }
  double upperBound(unsigned int) const { return m_Up; }
  double upperVelocityBound(unsigned int) const { return m_Vel; }
 private:
  double m_Lo, m_Up, m_Vel;
};
class CjrlFoot {
 public:
  CjrlFoot(double soleLength = 0.25, double soleWidth = 0.14, double ankleHeight = 0.105)
      : m_Length(soleLength), m_Width(soleWidth) { m_Ankle[0] = 0.0; m_Ankle[1] = 0.0; m_Ankle[2] = ankleHeight; }
  void getSoleSize(double &outLength, double &outWidth) const { outLength = m_Length; outWidth = m_Width; }
  void getAnklePositionInLocalFrame(vector3d &out) const { for (int i = 0; i < 3; ++i) out[i] = m_Ankle[i]; }
  CjrlJoint *associatedAnkle() { return &m_AnkleJoint; }
 private:
  double m_Length, m_Width, m_Ankle[3];
  CjrlJoint m_AnkleJoint;
};
/* Defaults: the sole of the reference's test robot (0.25 x 0.14 m) and hip-yaw joints without limits (equal
 * bounds make OrientationsPreview fall back to -30 / +45 degrees; a zero velocity bound is what the datref-era robot file
 * gave). */
class CjrlHumanoidDynamicRobot {
 public:
  CjrlHumanoidDynamicRobot(double soleLength = 0.25, double soleWidth = 0.14, double ankleHeight = 0.105)
      : m_Left(soleLength, soleWidth, ankleHeight), m_Right(soleLength, soleWidth, ankleHeight) {}
  CjrlFoot *leftFoot() { return &m_Left; }
  CjrlFoot *rightFoot() { return &m_Right; }
  CjrlJoint *waist() { return &m_Waist; }
  /* waist -> hip yaw -> ankle: element [1] is the hip-yaw joint of that leg, as the reference indexes it */
  std::vector<CjrlJoint *> jointsBetween(const CjrlJoint &, const CjrlJoint &inEndJoint)
  {
    const bool left = (&inEndJoint == m_Left.associatedAnkle());
    std::vector<CjrlJoint *> r;
    r.push_back(&m_Waist); r.push_back(left ? &m_LeftHipYaw : &m_RightHipYaw); r.push_back(left ? m_Left.associatedAnkle() : m_Right.associatedAnkle());
    return r;
  }
  void setHipYawJoints(const CjrlJoint &left, const CjrlJoint &right) { m_LeftHipYaw = left; m_RightHipYaw = right; }
  /* What ZMPPreviewControlWithMultiBodyZMP asks of the model (ZMPPreviewControlWithMultiBodyZMP.cpp:163-198, :447-479).  The
   * stand-in holds no multibody model: a caller with one overrides zeroMomentumPoint() (and positionCenterOfMass()). */
  virtual ~CjrlHumanoidDynamicRobot() {}
  const std::vector<double> &currentConfiguration() const { return m_Q; }
  const std::vector<double> &currentVelocity() const { return m_dQ; }
  const std::vector<double> &currentAcceleration() const { return m_ddQ; }
  bool currentConfiguration(const std::vector<double> &v) { m_Q = v; return true; }
  bool currentVelocity(const std::vector<double> &v) { m_dQ = v; return true; }
  bool currentAcceleration(const std::vector<double> &v) { m_ddQ = v; return true; }
  virtual bool setProperty(std::string &, const std::string &) { return true; }
  virtual bool getProperty(const std::string &, std::string &) { return true; }
  virtual bool computeForwardKinematics() { return true; }
  virtual vector3d zeroMomentumPoint() const { vector3d z; z[0] = z[1] = z[2] = 0.0; return z; }
  virtual vector3d positionCenterOfMass() const { vector3d z; z[0] = z[1] = z[2] = 0.0; return z; }
 private:
  CjrlFoot m_Left, m_Right;
  CjrlJoint m_Waist, m_LeftHipYaw, m_RightHipYaw;
  std::vector<double> m_Q, m_dQ, m_ddQ;
};
#endif

#define MAL_MATRIX(name, type) walkgen_b200::Matrix name
#define MAL_VECTOR_TYPE(type) std::vector<type>
#define MAL_MATRIX_DIM(name, type, r, c) walkgen_b200::Matrix name(r, c)
#define MAL_MATRIX_RESIZE(name, r, c) (name).resize(r, c)
#define MAL_MATRIX_NB_ROWS(name) (name).size1()
#define MAL_MATRIX_NB_COLS(name) (name).size2()

namespace PatternGeneratorJRL {

/* POD types of include/jrl/walkgen/pgtypes.hh (same field names and order). */
struct COMState {
  double x[3], y[3], z[3];
  double yaw[3], pitch[3], roll[3];
  COMState() { reset(); }
  void reset() { std::memset(this, 0, sizeof *this); }
};
struct ZMPPosition {
  double px, py, pz;
  double theta, time;
  int stepType;
};
struct FootAbsolutePosition {
  double x, y, z, theta, omega, omega2;
  double dx, dy, dz, dtheta, domega, domega2;
  double ddx, ddy, ddz, ddtheta, ddomega, ddomega2;
  double time;
  int stepType;
};

typedef COMState COMPosition;            /* include/jrl/walkgen/pgtypes.hh:88 (deprecated alias kept by the reference) */
struct RelativeFootPosition {            /* include/jrl/walkgen/pgtypes.hh:100-109 */
  double sx, sy, theta;
  double SStime, DStime;
  int stepType;
  double DeviationHipHeight;
};

class SimplePluginManager;

class SimplePlugin {
 public:
  explicit SimplePlugin(SimplePluginManager *lSPM) : m_SimplePluginManager(lSPM) {}
  virtual ~SimplePlugin();
  bool RegisterMethod(std::string &MethodName);
  virtual void CallMethod(std::string &Method, std::istringstream &astrm) = 0;
  SimplePluginManager *getSimplePluginManager() const { return m_SimplePluginManager; }
 private:
  SimplePluginManager *m_SimplePluginManager;
};

class SimplePluginManager {
 public:
  virtual ~SimplePluginManager() {}
  bool RegisterMethod(std::string &MethodName, SimplePlugin *aSP);
  void UnregisterPlugin(SimplePlugin *aSP);
  /* Broadcasts the rest of the buffer to EVERY plugin registered under the name (SimplePluginManager.cpp:107-162). */
  bool CallMethod(std::string &MethodName, std::istringstream &istrm);
 protected:
  std::multimap<std::string, SimplePlugin *> m_SimplePlugins;
};

struct OptimalControllerSolver {
  static const unsigned int MODE_WITH_INITIALPOS = 0;     /* WG_PREVIEW_MODE_WITH_INITIALPOS */
  static const unsigned int MODE_WITHOUT_INITIALPOS = 1;
};

class PreviewControl : public SimplePlugin {
 public:
  PreviewControl(SimplePluginManager *lSPM, unsigned int defaultMode = OptimalControllerSolver::MODE_WITH_INITIALPOS,
                 bool computeWeightsAutomatically = false);
  ~PreviewControl();
  /* Reads zc, T, preview time, Kx[3], Ks and the NL window weights; like the reference (PreviewControl.cpp:142-196)
   * every gain goes through a `float`, and an unreadable file only prints to cerr. */
  void ReadPrecomputedFile(std::string aFileName);
  /* x, y: 3 x 1 CoM state per axis (in/out).  Returns 0; throws std::runtime_error when fewer than the preview window
   * of ZMP positions is available from lindex on (the reference LTHROWs, PreviewControl.cpp:341-344). */
  int OneIterationOfPreview(MAL_MATRIX(&x, double), MAL_MATRIX(&y, double), double &sxzmp, double &syzmp,
                            std::deque<ZMPPosition> &ZMPPositions, unsigned int lindex, double &zmpx2, double &zmpy2,
                            bool Simulation);
  int OneIterationOfPreview1D(MAL_MATRIX(&x, double), double &sxzmp, std::deque<double> &ZMPPositions,
                              unsigned int lindex, double &zmpx2, bool Simulation);
  /* vector variant: the window wraps around the buffer (PreviewControl.cpp:448-466) */
  int OneIterationOfPreview1D(MAL_MATRIX(&x, double), double &sxzmp, std::vector<double> &ZMPPositions,
                              unsigned int lindex, double &zmpx2, bool Simulation);
  /* Batched form (new): every preview step of a whole ZMP reference in one call; com rows = (x,dx,ddx,y,dy,ddy). */
  int RunWholeTrajectory(const std::deque<ZMPPosition> &ZMPPositions, MAL_MATRIX(&x, double), MAL_MATRIX(&y, double),
                         double &sxzmp, double &syzmp, std::vector<double> &com6, std::vector<double> &zmp2,
                         bool Simulation);
  double SamplingPeriod() const { return m_SamplingPeriod; }
  double PreviewControlTime() const { return m_PreviewControlTime; }
  double GetHeightOfCoM() const { return m_Zc; }
  void SetSamplingPeriod(double lSamplingPeriod);
  void SetPreviewControlTime(double lPreviewControlTime);
  void SetHeightOfCoM(double lZc);
  bool IsCoherent() { return m_Coherent; }
  void ComputeOptimalWeights(unsigned int mode);
  void CallMethod(std::string &Method, std::istringstream &astrm);
  const wg_preview_gains_t &Gains() const { return m_Gains; }
  /* makes this object's gains the ones loaded in the process-wide context (done by every call above; public for the
   * classes that drive the batched C ABI with this controller's gains) */
  void BindGains() const;
 private:
  int run1d(walkgen_b200::Matrix &x, double &sxzmp, const std::vector<double> &window, double &zmpx2, bool Simulation);
  double m_SamplingPeriod, m_PreviewControlTime, m_Zc;
  bool m_Coherent, m_AutoComputeWeights;
  unsigned int m_DefaultWeightComputationMode;
  unsigned int m_SizeOfPreviewWindow;
  wg_preview_gains_t m_Gains;
};

class OptCholesky {
 public:
  OptCholesky(unsigned int lNbMaxOfConstraints, unsigned int lCardU, unsigned int mode);
  ~OptCholesky();
  void SetA(double *aA, unsigned int lNbOfConstraints);
  int AddActiveConstraints(std::vector<unsigned int> &lConstraints);
  int AddActiveConstraint(unsigned int aConstraint);
  int CurrentNumberOfRows();
  int ComputeNormalCholeskyOnANormal();
  int ComputeInverseCholeskyNormal(int mode);
  void SetL(double *aL);
  void SetiL(double *aiL);   /* unlike the reference (OptCholesky.cpp:116-121) a previously set iL is NOT deleted */
  void SetToZero();
  void SetMode(unsigned int mode) { m_UpdateMode = mode; }
  static const unsigned int MODE_NORMAL = 0;
  static const unsigned int MODE_FORTRAN = 1;
 private:
  unsigned int m_NbMaxOfConstraints, m_CardU;
  double *m_A, *m_L, *m_iL;
  unsigned int m_UpdateMode, m_NbOfConstraints;
  std::vector<unsigned int> m_SetActiveConstraints;
};

}  // namespace PatternGeneratorJRL

namespace Optimization {
namespace Solver {

class PLDPSolver {
 public:
  /* iLQ is accepted for signature parity; the reference only reads it in debug dumps. */
  PLDPSolver(unsigned int CardU, double *iPu, double *Px, double *Pu, double *iLQ);
  ~PLDPSolver();
  /* Returns 0, or -1 on NaN/Inf (PLDPSolver.cpp:955-964); -2 where the reference would exit(0) (negative step). */
  int SolveProblem(double *CstPartOfTheCostFunction, unsigned int NbOfConstraints, double *LinearPartOfConstraints,
                   double *CstPartOfConstraints, double *ZMPRef, double *XkYk, double *X,
                   std::vector<int> &SimilarConstraint, unsigned int NumberOfRemovedConstraints,
                   bool StartingSequence);
  const wg_pldp_info &LastInfo() const { return m_Info; }
 private:
  unsigned int m_CardV;
  wg_pldp_state m_Hot;
  wg_pldp_info m_Info;
};

}  // namespace Solver
}  // namespace Optimization

namespace PatternGeneratorJRL {

class StepStackHandler;

/* MAL_S3_VECTOR_TYPE(double) of the reference signatures: a 3-vector with (i) / [i] access */
struct S3Vector {
  double v[3];
  S3Vector() { v[0] = v[1] = v[2] = 0.0; }
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
#define MAL_S3_VECTOR_TYPE(type) PatternGeneratorJRL::S3Vector

/* The virtual interface of src/ZMPRefTrajectoryGeneration/ZMPRefTrajectoryGeneration.hh:208-328, signature for signature. */
class ZMPRefTrajectoryGeneration : public SimplePlugin {
 public:
  explicit ZMPRefTrajectoryGeneration(SimplePluginManager *lSPM);
  virtual ~ZMPRefTrajectoryGeneration() {}
  virtual void GetZMPDiscretization(std::deque<ZMPPosition> &ZMPPositions, std::deque<COMState> &COMStates,
                                    std::deque<RelativeFootPosition> &RelativeFootPositions,
                                    std::deque<FootAbsolutePosition> &LeftFootAbsolutePositions,
                                    std::deque<FootAbsolutePosition> &RightFootAbsolutePositions, double Xmax,
                                    COMState &lStartingCOMState, MAL_S3_VECTOR_TYPE(double) &lStartingZMPPosition,
                                    FootAbsolutePosition &InitLeftFootAbsolutePosition,
                                    FootAbsolutePosition &InitRightFootAbsolutePosition) = 0;
  virtual int InitOnLine(std::deque<ZMPPosition> &FinalZMPPositions, std::deque<COMState> &CoMStates,
                         std::deque<FootAbsolutePosition> &FinalLeftFootTraj_deq,
                         std::deque<FootAbsolutePosition> &FinalRightFootTraj_deq,
                         FootAbsolutePosition &InitLeftFootAbsolutePosition,
                         FootAbsolutePosition &InitRightFootAbsolutePosition,
                         std::deque<RelativeFootPosition> &RelativeFootPositions, COMState &lStartingCOMState,
                         MAL_S3_VECTOR_TYPE(double) &lStartingZMPPosition) = 0;
  virtual void OnLineAddFoot(RelativeFootPosition &NewRelativeFootPosition, std::deque<ZMPPosition> &FinalZMPPositions,
                             std::deque<COMState> &COMStates, std::deque<FootAbsolutePosition> &FinalLeftFootAbsolutePositions,
                             std::deque<FootAbsolutePosition> &FinalRightFootAbsolutePositions, bool EndSequence) = 0;
  virtual void OnLine(double time, std::deque<ZMPPosition> &FinalZMPPositions, std::deque<COMState> &CoMStates,
                      std::deque<FootAbsolutePosition> &FinalLeftFootTraj_deq,
                      std::deque<FootAbsolutePosition> &FinalRightFootTraj_deq) = 0;
  virtual void EndPhaseOfTheWalking(std::deque<ZMPPosition> &ZMPPositions, std::deque<COMState> &FinalCOMStates,
                                    std::deque<FootAbsolutePosition> &LeftFootAbsolutePositions,
                                    std::deque<FootAbsolutePosition> &RightFootAbsolutePositions) = 0;
  virtual int OnLineFootChange(double time, FootAbsolutePosition &aFootAbsolutePosition,
                               std::deque<ZMPPosition> &FinalZMPPositions, std::deque<COMState> &COMStates,
                               std::deque<FootAbsolutePosition> &FinalLeftFootAbsolutePositions,
                               std::deque<FootAbsolutePosition> &FinalRightFootAbsolutePositions,
                               StepStackHandler *aStepStackHandler) = 0;
  virtual int ReturnOptimalTimeToRegenerateAStep() = 0;
  virtual void CallMethod(std::string &Method, std::istringstream &strm);
  void SetTSingleSupport(double v) { m_Tsingle = v; }
  void SetTDoubleSupport(double v) { m_Tdble = v; }
  void SetSamplingPeriod(double v) { m_SamplingPeriod = v; }
  void SetComHeight(double v) { m_ComHeight = v; }
  double GetCurrentTime() const { return m_CurrentTime; }
  void SetCurrentTime(double t) { m_CurrentTime = t; }
  double GetTSingleSupport() const { return m_Tsingle; }
  double GetTDoubleSupport() const { return m_Tdble; }
  double GetSamplingPeriod() const { return m_SamplingPeriod; }
  double GetComHeight() const { return m_ComHeight; }
  bool GetOnLineMode() const { return m_OnLineMode; }
 protected:
  double m_Tsingle, m_Tdble, m_SamplingPeriod, m_Omega, m_ComHeight, m_StepHeight;
  double m_CurrentTime;
  bool m_OnLineMode;
};

/* Herdt's solution_t (src/privatepgtypes.hh:340-404): the fields a caller of ZMPVelocityReferencedQP::Solution() reads. */
struct support_state_t {                 /* src/privatepgtypes.hh:291-320 */
  int Phase, Foot;                       /* WG_SS / WG_DS, WG_LEFT / WG_RIGHT */
  unsigned int StepNumber;
  bool StateChanged;
  double X, Y, Yaw;
};
struct solution_t {
  unsigned int NbVariables, NbConstraints;
  int Fail, Print;
  bool useWarmStart;
  std::vector<double> Solution_vec, initialSolution;
  std::deque<double> SupportOrientations_deq, TrunkOrientations_deq;
  std::deque<support_state_t> SupportStates_deq;
  std::vector<double> ConstrLagr_vec, LBoundsLagr_vec, UBoundsLagr_vec;
  solution_t() : NbVariables(0), NbConstraints(0), Fail(0), Print(0), useWarmStart(false) {}
};

class ZMPVelocityReferencedQP : public ZMPRefTrajectoryGeneration {
 public:
  /* the robot is only asked for its sole size in the reference (RelativeFeetInequalities): pass it directly */
  ZMPVelocityReferencedQP(SimplePluginManager *lSPM, std::string DataFile, double sole_length = 0.25,
                          double sole_width = 0.14);
  /* the reference's constructor (ZMPVelocityReferencedQP.hh:59-60): sole size and hip-yaw joint bounds are read from the
   * robot as RelativeFeetInequalities / OrientationsPreview do */
  ZMPVelocityReferencedQP(SimplePluginManager *lSPM, std::string DataFile, CjrlHumanoidDynamicRobot *aHS);
  ~ZMPVelocityReferencedQP();
  int InitOnLine(std::deque<ZMPPosition> &FinalZMPPositions, std::deque<COMState> &CoMStates,
                 std::deque<FootAbsolutePosition> &FinalLeftFootTraj_deq,
                 std::deque<FootAbsolutePosition> &FinalRightFootTraj_deq,
                 FootAbsolutePosition &InitLeftFootAbsolutePosition, FootAbsolutePosition &InitRightFootAbsolutePosition,
                 std::deque<RelativeFootPosition> &RelativeFootPositions, COMState &lStartingCOMState,
                 MAL_S3_VECTOR_TYPE(double) &lStartingZMPPosition);
  /* the off-line / foot-by-foot entry points are empty in the reference too (ZMPVelocityReferencedQP.cpp:462-521) */
  void GetZMPDiscretization(std::deque<ZMPPosition> &, std::deque<COMState> &, std::deque<RelativeFootPosition> &,
                            std::deque<FootAbsolutePosition> &, std::deque<FootAbsolutePosition> &, double, COMState &,
                            MAL_S3_VECTOR_TYPE(double) &, FootAbsolutePosition &, FootAbsolutePosition &) {}
  void OnLineAddFoot(RelativeFootPosition &, std::deque<ZMPPosition> &, std::deque<COMState> &,
                     std::deque<FootAbsolutePosition> &, std::deque<FootAbsolutePosition> &, bool) {}
  void EndPhaseOfTheWalking(std::deque<ZMPPosition> &, std::deque<COMState> &, std::deque<FootAbsolutePosition> &,
                            std::deque<FootAbsolutePosition> &) {}
  int OnLineFootChange(double, FootAbsolutePosition &, std::deque<ZMPPosition> &, std::deque<COMState> &,
                       std::deque<FootAbsolutePosition> &, std::deque<FootAbsolutePosition> &, StepStackHandler *) { return -1; }
  int ReturnOptimalTimeToRegenerateAStep() { return 2 * (int)(1.6 / m_SamplingPeriod); }
  /* Solution_ of the last QP (ZMPVelocityReferencedQP.hh:126): Solution_vec, multipliers and the previewed support states,
   * obtained by solving the recorded QP of the last period once more through wg_herdt_qp_solve_batch */
  solution_t &Solution();
  void OnLine(double time, std::deque<ZMPPosition> &FinalZMPPositions, std::deque<COMState> &CoMStates,
              std::deque<FootAbsolutePosition> &FinalLeftFootTraj_deq,
              std::deque<FootAbsolutePosition> &FinalRightFootTraj_deq);
  void Reference(std::istringstream &strm) { strm >> m_State.new_ref[0] >> m_State.new_ref[1] >> m_State.new_ref[2]; }
  void Reference(double dx, double dy, double dyaw) { m_State.new_ref[0] = dx; m_State.new_ref[1] = dy; m_State.new_ref[2] = dyaw; }
  bool Running() const { return m_State.running != 0; }
  void EndingPhase(bool EndingPhase) { m_State.ending_phase = EndingPhase; }
  void setCoMPerturbationForce(double, double) {}   /* parsed but never consumed by the reference either */
  void setCoMPerturbationForce(std::istringstream &strm) { double x, y; strm >> x >> y; }
  unsigned QP_N() const { return WG_HERDT_N; }
  void CallMethod(std::string &Method, std::istringstream &strm);
  /* datref-era initial support frame ("oracle pins") */
  void SetInitialSupportFrame(double x, double y, double yaw) { m_State.sup_x = x; m_State.sup_y = y; m_State.sup_yaw = yaw; }
  /* what OrientationsPreview's ctor reads from the CjrlHumanoidDynamicRobot (OrientationsPreview.cpp:48-68): hip-yaw
   * lower/upper bounds (equal bounds -> the reference's -30/+45 deg defaults) and |upperVelocityBound| */
  void SetHipYawJoints(double lLeft, double uLeft, double lRight, double uRight, double upperVelocityBound);
  /* end-of-walk jerk towards the feet centre (ZMPVelocityReferencedQP.cpp:410-421, since 3.1.8); default on */
  void SetReturnToCentre(bool on) { m_Params.return_to_centre = on; m_ParamsDirty = true; }
  /* the three settings under which the reference's committed TestHerdt2010 datrefs are reproduced */
  void SetDatrefEra() { SetInitialSupportFrame(0.0, 0.1, 0.0); SetHipYawJoints(0, 0, 0, 0, 0.0); SetReturnToCentre(false); }
  const wg_herdt_mpc_state &State() const { return m_State; }
 private:
  wg_herdt_mpc_state m_State;
  wg_herdt_mpc_params m_Params;
  double m_SoleLength, m_SoleWidth;
  bool m_ParamsDirty;
  unsigned m_StepsBeforeStop;
  wg_herdt_qp_input m_LastQP;
  bool m_HaveLastQP;
  solution_t m_Solution;
};

/* ---- Kajita2003 front end: the step stack and the ZMP / feet discretisation -------------------------------------- */
class StepStackHandler : public SimplePlugin {   /* src/StepStackHandler.hh */
 public:
  explicit StepStackHandler(SimplePluginManager *lSPM);
  void ReadStepSequenceAccordingToWalkMode(std::istringstream &strm);   /* walk modes 0, 4, 5 (1 / 3 read and drop the hip height) */
  void CreateArcInStepStack(double x, double y, double R, double arc_deg, int SupportFoot);
  void CreateArcCenteredInStepStack(double R, double arc_deg, int SupportFoot);   /* not provided: throws */
  void PrepareForSupportFoot(int SupportFoot);
  void FinishOnTheLastCorrectSupportFoot();
  void AddStepInTheStack(double sx, double sy, double theta, double sstime, double dstime);
  void AddStandardOnLineStep(bool NewStep, double NewStepX, double NewStepY, double NewTheta);
  void PushFrontAStepInTheStack(RelativeFootPosition &aRFP) { m_RelativeFootPositions.push_front(aRFP); }
  bool RemoveFirstStepInTheStack();
  void CopyRelativeFootPosition(std::deque<RelativeFootPosition> &lRelativeFootPositions, bool PerformClean);
  RelativeFootPosition ReturnBackFootPosition() { return m_RelativeFootPositions.back(); }
  bool ReturnFrontFootPosition(RelativeFootPosition &aRFP);
  int ReturnStackSize() { return (int)m_RelativeFootPositions.size(); }
  void ClearStack() { m_RelativeFootPositions.clear(); }
  void SetWalkMode(int lWalkMode) { m_WalkMode = lWalkMode; }
  int GetWalkMode() { return m_WalkMode; }
  void SetSingleTimeSupport(double v) { m_SingleSupportTime = v; }
  double GetSingleTimeSupport() { return m_SingleSupportTime; }
  void SetDoubleTimeSupport(double v) { m_DoubleSupportTime = v; }
  double GetDoubleTimeSupport() { return m_DoubleSupportTime; }
  void StartOnLineStep() { m_OnLineSteps = true; }
  void StopOnLineStep();
  bool IsOnLineSteppingOn() { return m_OnLineSteps; }
  void m_PartialStepSequence(std::istringstream &strm);
  void CallMethod(std::string &Method, std::istringstream &strm);
 private:
  std::deque<RelativeFootPosition> m_RelativeFootPositions;
  double m_SingleSupportTime, m_DoubleSupportTime;
  int m_WalkMode, m_KeepLastCorrectSupportFoot;
  bool m_OnLineSteps, m_TransitionFinishOnLine;
};

/* ZMPDiscretization (src/ZMPRefTrajectoryGeneration/ZMPDiscretization.hh): footsteps -> 5 ms ZMP reference + both feet.
 * Every sample comes from zmpdisc_kernel (wg_zmpdisc_run_batch, batch of one walk).  The on-line entry points rest on
 * the prefix property of the generator (a walk is a chain of segments - lead-in, one per step, end phase - each of which
 * only depends on the state the previous one left): the walk given so far is discretised again and the samples not yet
 * handed out are appended to the caller's queues. */
class ZMPDiscretization : public ZMPRefTrajectoryGeneration {
 public:
  ZMPDiscretization(SimplePluginManager *lSPM, std::string DataFile = "", CjrlHumanoidDynamicRobot *aHS = 0);
  ~ZMPDiscretization();
  void GetZMPDiscretization(std::deque<ZMPPosition> &ZMPPositions, std::deque<COMState> &COMStates,
                            std::deque<RelativeFootPosition> &RelativeFootPositions,
                            std::deque<FootAbsolutePosition> &LeftFootAbsolutePositions,
                            std::deque<FootAbsolutePosition> &RightFootAbsolutePositions, double Xmax,
                            COMState &lStartingCOMState, MAL_S3_VECTOR_TYPE(double) &lStartingZMPPosition,
                            FootAbsolutePosition &InitLeftFootAbsolutePosition,
                            FootAbsolutePosition &InitRightFootAbsolutePosition);
  int InitOnLine(std::deque<ZMPPosition> &FinalZMPPositions, std::deque<COMState> &CoMStates,
                 std::deque<FootAbsolutePosition> &FinalLeftFootAbsolutePositions,
                 std::deque<FootAbsolutePosition> &FinalRightFootAbsolutePositions,
                 FootAbsolutePosition &InitLeftFootAbsolutePosition, FootAbsolutePosition &InitRightFootAbsolutePosition,
                 std::deque<RelativeFootPosition> &RelativeFootPositions, COMState &lStartingCOMState,
                 MAL_S3_VECTOR_TYPE(double) &lStartingZMPPosition);
  void OnLineAddFoot(RelativeFootPosition &NewRelativeFootPosition, std::deque<ZMPPosition> &FinalZMPPositions,
                     std::deque<COMState> &COMStates, std::deque<FootAbsolutePosition> &FinalLeftFootAbsolutePositions,
                     std::deque<FootAbsolutePosition> &FinalRightFootAbsolutePositions, bool EndSequence);
  /* the Kajita generator has no per-tick work: its queues are filled foot by foot (ZMPDiscretization.cpp:562-571) */
  void OnLine(double, std::deque<ZMPPosition> &, std::deque<COMState> &, std::deque<FootAbsolutePosition> &,
              std::deque<FootAbsolutePosition> &) {}
  void EndPhaseOfTheWalking(std::deque<ZMPPosition> &ZMPPositions, std::deque<COMState> &FinalCOMStates,
                            std::deque<FootAbsolutePosition> &LeftFootAbsolutePositions,
                            std::deque<FootAbsolutePosition> &RightFootAbsolutePositions);
  /* returns -1 in the reference as well (ZMPDiscretization.cpp:1111-1120) */
  int OnLineFootChange(double, FootAbsolutePosition &, std::deque<ZMPPosition> &, std::deque<COMState> &,
                       std::deque<FootAbsolutePosition> &, std::deque<FootAbsolutePosition> &, StepStackHandler *) { return -1; }
  int ReturnOptimalTimeToRegenerateAStep();
  void SetZMPShift(std::vector<double> &ZMPShift);
  void SetPreviewControlTime(double v) { m_PreviewControlTime = v; }
  void CallMethod(std::string &Method, std::istringstream &strm);
 private:
  /* discretise m_Steps (+ end phase) and append samples [m_Emitted, upto) to the queues */
  void emit(std::deque<ZMPPosition> &Z, std::deque<COMState> &C, std::deque<FootAbsolutePosition> &L,
            std::deque<FootAbsolutePosition> &R, bool with_end_phase);
  wg_zmpdisc_params m_Zd;
  double m_PreviewControlTime;
  std::vector<wg_rel_step> m_Steps;
  double m_InitFeet[6];
  int64_t m_Emitted;
};

/* ---- Dimitrov2008 path: the classes around PLDPSolver (names and signatures of the reference) ------------------------ */
typedef struct { double col, row; } CH_Point;                       /* src/Mathematics/ConvexHull.hh:39-42 */
struct LinearConstraintInequality_t {    /* include/jrl/walkgen/pgtypes.hh:168-177; A z + B >= 0 */
  MAL_MATRIX(A, double);
  MAL_MATRIX(B, double);
  std::vector<double> Center;
  std::vector<int> SimilarConstraints;
  double StartingTime, EndingTime;
};

class ComputeConvexHull {                /* src/Mathematics/ConvexHull.hh:47-60 */
 public:
  void DoComputeConvexHull(std::vector<CH_Point> aVecOfPoints, std::vector<CH_Point> &TheConvexHull);
};

class FootConstraintsAsLinearSystem : public SimplePlugin {   /* src/Mathematics/FootConstraintsAsLinearSystem.hh:54-120 */
 public:
  /* the robot is only asked for the sole size of its feet (FootConstraintsAsLinearSystem.cpp:269-281): pass it directly */
  FootConstraintsAsLinearSystem(SimplePluginManager *aSPM, double sole_length = 0.25, double sole_width = 0.14);
  int BuildLinearConstraintInequalities(std::deque<FootAbsolutePosition> &LeftFootAbsolutePositions,
                                        std::deque<FootAbsolutePosition> &RightFootAbsolutePositions,
                                        std::deque<LinearConstraintInequality_t *> &QueueOfLConstraintInequalities,
                                        double ConstraintOnX, double ConstraintOnY);
  void CallMethod(std::string &, std::istringstream &) {}
 private:
  double m_SoleLength, m_SoleWidth;
};

class ZMPConstrainedQPFastFormulation : public ZMPRefTrajectoryGeneration {   /* ...ZMPConstrainedQPFastFormulation.hh */
 public:
  static const unsigned int PLDP = 2;
  ZMPConstrainedQPFastFormulation(SimplePluginManager *lSPM, std::string DataFile, double sole_length = 0.25,
                                  double sole_width = 0.14);
  /* the whole walk, off line, as in the reference: ZMPDiscretization + BuildZMPTrajectoryFromFootTrajectory (PLDP).
   * Returns through the deques; LastStatus() = 0, or 1 where the reference prints IFAIL / calls exit(0). */
  void GetZMPDiscretization(std::deque<ZMPPosition> &ZMPPositions, std::deque<COMState> &COMStates,
                            std::deque<RelativeFootPosition> &RelativeFootPositions,
                            std::deque<FootAbsolutePosition> &LeftFootAbsolutePositions,
                            std::deque<FootAbsolutePosition> &RightFootAbsolutePositions, double Xmax,
                            COMState &lStartingCOMState, MAL_S3_VECTOR_TYPE(double) &lStartingZMPPosition,
                            FootAbsolutePosition &InitLeftFootAbsolutePosition,
                            FootAbsolutePosition &InitRightFootAbsolutePosition);
  int InitConstants();
  void SetAlpha(const double &a) { m_Par.alpha = a; m_Dirty = true; }
  const double &GetAlpha() const { return m_Par.alpha; }
  void SetBeta(const double &b) { m_Par.beta = b; m_Dirty = true; }
  const double &GetBeta() const { return m_Par.beta; }
  /* :setdimitrovconstraint X Y; the ZMPRefTrajectoryGeneration commands reach the embedded ZMPDiscretization */
  void CallMethod(std::string &Method, std::istringstream &strm);
  /* not in the reference (both off by default): wg_dimitrov_params.cold_restart / merge_duplicate_rows */
  void SetRobustMode(bool cold_restart, bool merge_duplicate_rows)
  { m_Par.cold_restart = cold_restart; m_Par.merge_duplicate_rows = merge_duplicate_rows; m_Dirty = true; }
  int LastStatus() const { return m_Status; }
  int PeriodsDone() const { return m_Done; }
  /* InitOnLine / OnLine are not provided by the reference for this generator either (they return 0 / do nothing) */
  int InitOnLine(std::deque<ZMPPosition> &, std::deque<COMState> &, std::deque<FootAbsolutePosition> &,
                 std::deque<FootAbsolutePosition> &, FootAbsolutePosition &, FootAbsolutePosition &,
                 std::deque<RelativeFootPosition> &, COMState &, MAL_S3_VECTOR_TYPE(double) &) { return 0; }
  void OnLineAddFoot(RelativeFootPosition &, std::deque<ZMPPosition> &, std::deque<COMState> &,
                     std::deque<FootAbsolutePosition> &, std::deque<FootAbsolutePosition> &, bool) {}
  void OnLine(double, std::deque<ZMPPosition> &, std::deque<COMState> &, std::deque<FootAbsolutePosition> &,
              std::deque<FootAbsolutePosition> &) {}
  void EndPhaseOfTheWalking(std::deque<ZMPPosition> &, std::deque<COMState> &, std::deque<FootAbsolutePosition> &,
                            std::deque<FootAbsolutePosition> &) {}
  int OnLineFootChange(double, FootAbsolutePosition &, std::deque<ZMPPosition> &, std::deque<COMState> &,
                       std::deque<FootAbsolutePosition> &, std::deque<FootAbsolutePosition> &, StepStackHandler *) { return -1; }
  int ReturnOptimalTimeToRegenerateAStep() { return 0; }
 private:
  wg_dimitrov_params m_Par;
  wg_zmpdisc_params m_Zd;
  bool m_Dirty;
  int m_Status, m_Done;
};

/* src/ZMPRefTrajectoryGeneration/ZMPQPWithConstraint.hh: the Wieber2006 generator (the whole walk off line, as in the
 * reference: ZMPDiscretization, then one 150-variable QP per 20 ms).  ":setpbwconstraint XY x y | T t | N n" as :1389-1414. */
class ZMPQPWithConstraint : public ZMPRefTrajectoryGeneration {
 public:
  ZMPQPWithConstraint(SimplePluginManager *lSPM, std::string DataFile, CjrlHumanoidDynamicRobot *aHS = 0);
  void GetZMPDiscretization(std::deque<ZMPPosition> &ZMPPositions, std::deque<COMState> &COMStates,
                            std::deque<RelativeFootPosition> &RelativeFootPositions,
                            std::deque<FootAbsolutePosition> &LeftFootAbsolutePositions,
                            std::deque<FootAbsolutePosition> &RightFootAbsolutePositions, double Xmax,
                            COMState &lStartingCOMState, MAL_S3_VECTOR_TYPE(double) &lStartingZMPPosition,
                            FootAbsolutePosition &InitLeftFootAbsolutePosition,
                            FootAbsolutePosition &InitRightFootAbsolutePosition);
  void CallMethod(std::string &Method, std::istringstream &strm);
  /* 0, or the wg_wieber_run_batch status where the reference prints and returns -1 */
  int LastStatus() const { return m_Status; }
  int PeriodsDone() const { return m_Done; }
  /* 0: solve the QPs as stated instead of reproducing ql0001_'s regularised Hessian (wg_wieber_params::qld_eps) */
  void SetQLDEps(double eps) { m_Par.qld_eps = eps; }
  /* "To be implemented" in the reference as well (:1417-1474) */
  int InitOnLine(std::deque<ZMPPosition> &, std::deque<COMState> &, std::deque<FootAbsolutePosition> &,
                 std::deque<FootAbsolutePosition> &, FootAbsolutePosition &, FootAbsolutePosition &,
                 std::deque<RelativeFootPosition> &, COMState &, MAL_S3_VECTOR_TYPE(double) &) { return 0; }
  void OnLineAddFoot(RelativeFootPosition &, std::deque<ZMPPosition> &, std::deque<COMState> &,
                     std::deque<FootAbsolutePosition> &, std::deque<FootAbsolutePosition> &, bool) {}
  void OnLine(double, std::deque<ZMPPosition> &, std::deque<COMState> &, std::deque<FootAbsolutePosition> &,
              std::deque<FootAbsolutePosition> &) {}
  void EndPhaseOfTheWalking(std::deque<ZMPPosition> &, std::deque<COMState> &, std::deque<FootAbsolutePosition> &,
                            std::deque<FootAbsolutePosition> &) {}
  int OnLineFootChange(double, FootAbsolutePosition &, std::deque<ZMPPosition> &, std::deque<COMState> &,
                       std::deque<FootAbsolutePosition> &, std::deque<FootAbsolutePosition> &, StepStackHandler *) { return -1; }
  int ReturnOptimalTimeToRegenerateAStep() { return 0; }
 private:
  wg_wieber_params m_Par;
  wg_zmpdisc_params m_Zd;
  int m_Status, m_Done;
};

/* The PGI facade (include/jrl/walkgen/patterngeneratorinterface.hh:55-306) for the accelerated paths: command bus + the
 * 5 ms tick, for the Herdt on-line generator and for the Kajita off-line / on-line step sequences.  Whole-body inverse
 * kinematics and the multibody second preview stage are outside the accelerated path: as with the
 * reference's CoMAndFootOnlyStrategy the configuration / velocity / acceleration vectors are left as the caller passed them. */
class PatternGeneratorInterface : public SimplePluginManager, public SimplePlugin {
 public:
  PatternGeneratorInterface(double sole_length = 0.25, double sole_width = 0.14);
  explicit PatternGeneratorInterface(CjrlHumanoidDynamicRobot *aHDR);
  ~PatternGeneratorInterface();
  /* Reads the first token and broadcasts the rest (PatternGeneratorInterfacePrivate.cpp:1030-1041). */
  int ParseCmd(std::istringstream &strm);
  void CallMethod(std::string &Method, std::istringstream &strm);
  /* One 5 ms tick (PatternGeneratorInterfacePrivate.cpp:1246-1336, Herdt branch + CoMAndFootOnlyStrategy pop).
   * Returns false when the deques ran empty (end of the motion). */
  bool RunOneStepOfTheControlLoop(COMState &COMStateOut, ZMPPosition &ZMPTarget, FootAbsolutePosition &LeftFootPosition,
                                  FootAbsolutePosition &RightFootPosition);
  /* the four overloads of patterngeneratorinterface.hh:115-176 */
  bool RunOneStepOfTheControlLoop(MAL_VECTOR_TYPE(double) &CurrentConfiguration, MAL_VECTOR_TYPE(double) &CurrentVelocity,
                                  MAL_VECTOR_TYPE(double) &CurrentAcceleration, MAL_VECTOR_TYPE(double) &ZMPTarget);
  bool RunOneStepOfTheControlLoop(MAL_VECTOR_TYPE(double) &CurrentConfiguration, MAL_VECTOR_TYPE(double) &CurrentVelocity,
                                  MAL_VECTOR_TYPE(double) &CurrentAcceleration, MAL_VECTOR_TYPE(double) &ZMPTarget,
                                  COMState &COMState, FootAbsolutePosition &LeftFootPosition,
                                  FootAbsolutePosition &RightFootPosition);
  bool RunOneStepOfTheControlLoop(FootAbsolutePosition &LeftFootPosition, FootAbsolutePosition &RightFootPosition,
                                  ZMPPosition &ZMPRefPos, COMPosition &COMRefPos);
  /* (the COMPosition flavour of the 7-argument overload is the same function: COMPosition is a typedef of COMState) */
  void ReadSequenceOfSteps(std::istringstream &strm);
  void FinishAndRealizeStepSequence();
  void StartOnLineStepSequencing();
  void StopOnLineStepSequencing();
  void AddOnLineStep(double X, double Y, double Theta);
  void AddStepInStack(double dx, double dy, double theta);
  int GetWalkMode() const { return m_StepStackHandler->GetWalkMode(); }
  void setVelocityReference(double x, double y, double yaw) { m_ZMPVRQP->Reference(x, y, yaw); }
  StepStackHandler *SSH() { return m_StepStackHandler; }
  ZMPDiscretization *ZMPD() { return m_ZMPD; }
  /* CoM of the first preview stage for the tick last returned (the Kajita path; the reference's second stage needs the
   * multibody robot model) */
  const std::deque<COMState> &COMBuffer() const { return m_COMBuffer; }
  void SetStartConfiguration(const COMState &com, const FootAbsolutePosition &lf, const FootAbsolutePosition &rf);
  ZMPVelocityReferencedQP *VRQP() { return m_ZMPVRQP; }
  PreviewControl *PC() { return m_PC; }
 private:
  int initOnlineHerdt();
  void construct(double sole_length, double sole_width, CjrlHumanoidDynamicRobot *aHDR);
  void kajitaPreviewOverQueues();
  ZMPVelocityReferencedQP *m_ZMPVRQP;
  PreviewControl *m_PC;
  StepStackHandler *m_StepStackHandler;
  ZMPDiscretization *m_ZMPD;
  CjrlHumanoidDynamicRobot *m_OwnRobot;
  std::vector<double> m_ZMPShift;
  bool m_AutoFirstStep, m_KajitaOnLine;
  double m_PreviewState[8];             /* x[3], y[3], sxzmp, syzmp of the first preview stage */
  size_t m_PreviewedUpTo;               /* samples of the queues whose CoM has been computed */
  std::deque<ZMPPosition> m_ZMPPositions;
  std::deque<COMState> m_COMBuffer;
  std::deque<FootAbsolutePosition> m_LeftFootPositions, m_RightFootPositions;
  COMState m_StartCOM;
  FootAbsolutePosition m_StartLF, m_StartRF;
  double m_InternalClock;
  int m_AlgorithmforZMPCOM;   /* 0 Kajita (default), 1 Herdt */
  bool m_Running;
};

/* The interface of src/MotionGeneration/ComAndFootRealization.hh:55-219 that the two-stage scheme calls: whole-body
 * realisation of a CoM + feet posture (IK in the reference; out of this library's scope - the caller supplies it). */
class ComAndFootRealization {
 public:
  ComAndFootRealization() : m_HumanoidDynamicRobot(0) {}
  virtual ~ComAndFootRealization() {}
  virtual bool ComputePostureForGivenCoMAndFeetPosture(MAL_VECTOR_TYPE(double) &CoMPosition, MAL_VECTOR_TYPE(double) &CoMSpeed,
                                                       MAL_VECTOR_TYPE(double) &CoMAcc, MAL_VECTOR_TYPE(double) &LeftFoot,
                                                       MAL_VECTOR_TYPE(double) &RightFoot,
                                                       MAL_VECTOR_TYPE(double) &CurrentConfiguration,
                                                       MAL_VECTOR_TYPE(double) &CurrentVelocity,
                                                       MAL_VECTOR_TYPE(double) &CurrentAcceleration, int IterationNumber,
                                                       int Stage) = 0;
  virtual bool InitializationCoM(MAL_VECTOR_TYPE(double) &BodyAnglesIni, MAL_S3_VECTOR_TYPE(double) &lStartingCOMPosition,
                                 MAL_VECTOR_TYPE(double) &lStartingWaistPose, FootAbsolutePosition &InitLeftFootAbsPos,
                                 FootAbsolutePosition &InitRightFootAbsPos) = 0;
  virtual MAL_S3_VECTOR_TYPE(double) GetCOGInitialAnkles() { return MAL_S3_VECTOR_TYPE(double)(); }
  virtual bool setHumanoidDynamicRobot(CjrlHumanoidDynamicRobot *aHumanoidDynamicRobot)
  { m_HumanoidDynamicRobot = aHumanoidDynamicRobot; return true; }
  CjrlHumanoidDynamicRobot *getHumanoidDynamicRobot() const { return m_HumanoidDynamicRobot; }
 private:
  CjrlHumanoidDynamicRobot *m_HumanoidDynamicRobot;
};

/* src/PreviewControl/ZMPPreviewControlWithMultiBodyZMP.hh: Kajita's two-stage scheme.  The FIFO bookkeeping is the
 * reference's, statement for statement (Setup skipping ZMPRefPositions[NL] included); each stage's
 * OneIterationOfPreview goes to the GPU through the PreviewControl mirror.  The multibody ZMP comes from the robot
 * (zeroMomentumPoint(), as in the reference), the posture from the ComAndFootRealization: both are the caller's. */
class ZMPPreviewControlWithMultiBodyZMP : public SimplePlugin {
 public:
  static const int ZMPCOM_TRAJECTORY_FULL = 1;
  static const int ZMPCOM_TRAJECTORY_SECOND_STAGE_ONLY = 2;
  static const int ZMPCOM_TRAJECTORY_FIRST_STAGE_ONLY = 3;
  explicit ZMPPreviewControlWithMultiBodyZMP(SimplePluginManager *lSPM);
  ~ZMPPreviewControlWithMultiBodyZMP();
  void SetPreviewControl(PreviewControl *aPC);
  void SetStrategyForStageActivation(int aZMPComTraj);
  int GetStrategyForStageActivation() { return m_StageStrategy; }
  void SetStrategyForPCStages(int Strategy) { m_StageStrategy = Strategy; }
  int GetStrategyForPCStages() { return m_StageStrategy; }
  bool setComAndFootRealization(ComAndFootRealization *aCFR) { m_ComAndFootRealization = aCFR; return true; }
  ComAndFootRealization *getComAndFootRealization() { return m_ComAndFootRealization; }
  bool setHumanoidDynamicRobot(CjrlHumanoidDynamicRobot *aHumanoidDynamicRobot)
  { m_HumanoidDynamicRobot = aHumanoidDynamicRobot; return true; }
  CjrlHumanoidDynamicRobot *getHumanoidDynamicRobot() const { return m_HumanoidDynamicRobot; }
  int OneGlobalStepOfControl(FootAbsolutePosition &LeftFootPosition, FootAbsolutePosition &RightFootPosition,
                             ZMPPosition &NewZMPRefPos, COMState &refandfinalCOMState,
                             MAL_VECTOR_TYPE(double) &CurrentConfiguration, MAL_VECTOR_TYPE(double) &CurrentVelocity,
                             MAL_VECTOR_TYPE(double) &CurrentAcceleration);
  int FirstStageOfControl(FootAbsolutePosition &LeftFootPosition, FootAbsolutePosition &RightFootPosition,
                          COMState &afCOMState);
  int EvaluateMultiBodyZMP(int StartingIteration);
  int SecondStageOfControl(COMState &refandfinalCOMState);
  COMState GetLastCOMFromFirstStage() { return m_FIFOCOMStates.back(); }
  int Setup(std::deque<ZMPPosition> &ZMPRefPositions, std::deque<COMState> &COMStates,
            std::deque<FootAbsolutePosition> &LeftFootPositions, std::deque<FootAbsolutePosition> &RightFootPositions);
  int SetupFirstPhase(std::deque<ZMPPosition> &ZMPRefPositions, std::deque<COMState> &COMStates,
                      std::deque<FootAbsolutePosition> &LeftFootPositions,
                      std::deque<FootAbsolutePosition> &RightFootPositions);
  int SetupIterativePhase(std::deque<ZMPPosition> &ZMPRefPositions, std::deque<COMState> &COMStates,
                          std::deque<FootAbsolutePosition> &LeftFootPositions,
                          std::deque<FootAbsolutePosition> &RightFootPositions,
                          MAL_VECTOR_TYPE(double) &CurrentConfiguration, MAL_VECTOR_TYPE(double) &CurrentVelocity,
                          MAL_VECTOR_TYPE(double) &CurrentAcceleration, int localindex);
  void CreateExtraCOMBuffer(std::deque<COMState> &ExtraCOMBuffer, std::deque<ZMPPosition> &ExtraZMPBuffer,
                            std::deque<ZMPPosition> &ExtraZMPRefBuffer);
  int EvaluateStartingCoM(MAL_VECTOR_TYPE(double) &BodyAnglesInit, MAL_S3_VECTOR_TYPE(double) &aStartingCOMState,
                          MAL_VECTOR_TYPE(double) &aStartingWaistPosition, FootAbsolutePosition &InitLeftFootPosition,
                          FootAbsolutePosition &InitRightFootPosition);
  int EvaluateStartingState(MAL_VECTOR_TYPE(double) &BodyAnglesInit, MAL_S3_VECTOR_TYPE(double) &aStartingCOMState,
                            MAL_S3_VECTOR_TYPE(double) &aStartingZMPPosition, MAL_VECTOR_TYPE(double) &aStartingWaistPosition,
                            FootAbsolutePosition &InitLeftFootPosition, FootAbsolutePosition &InitRightFootPosition);
  void UpdateTheZMPRefQueue(ZMPPosition NewZMPRefPos) { m_FIFOZMPRefPositions.push_back(NewZMPRefPos); }
  void SetSamplingPeriod(double v) { m_SamplingPeriod = v; }
  double SamplingPeriod() const { return m_SamplingPeriod; }
  void SetPreviewControlTime(double v) { m_PreviewControlTime = v; }
  double PreviewControlTime() const { return m_PreviewControlTime; }
  void CallMethod(std::string &Method, std::istringstream &astrm);
  /* Batched form (new): the whole scheme over a complete ZMP reference with the multibody ZMP of every first-stage tick
   * supplied by `multibody_zmp(tick, CoM of that tick, out xy)`: one GPU pass per stage instead of two launches per tick.
   * Same FIFO semantics as Setup + OneGlobalStepOfControl to the end of the stream; FinalCOMStates gets one state per
   * global step (x, y filled). */
  int RunWholeTrajectory(std::deque<ZMPPosition> &ZMPRefPositions, const COMState &StartingCOM,
                         void (*multibody_zmp)(void *user, long tick, const double *com6, double *zmp_xy), void *user,
                         std::deque<COMState> &FinalCOMStates);
 private:
  void CallToComAndFootRealization(COMState &acomp, FootAbsolutePosition &aLeftFAP, FootAbsolutePosition &aRightFAP,
                                   MAL_VECTOR_TYPE(double) &CurrentConfiguration, MAL_VECTOR_TYPE(double) &CurrentVelocity,
                                   MAL_VECTOR_TYPE(double) &CurrentAcceleration, int IterationNumber, int StageOfTheAlgorithm);
  PreviewControl *m_PC;
  bool m_OwnPC;
  CjrlHumanoidDynamicRobot *m_HumanoidDynamicRobot;
  ComAndFootRealization *m_ComAndFootRealization;
  walkgen_b200::Matrix m_PC1x, m_PC1y, m_Deltax, m_Deltay;
  double m_sxzmp, m_syzmp, m_sxDeltazmp, m_syDeltazmp;
  double m_SamplingPeriod, m_PreviewControlTime;
  unsigned int m_NL;
  int m_StageStrategy, m_NumberOfIterations;
  bool m_StartingNewSequence;
  MAL_S3_VECTOR_TYPE(double) m_StartingCOMState;
  std::deque<ZMPPosition> m_FIFOZMPRefPositions, m_FIFODeltaZMPPositions;
  std::deque<COMState> m_FIFOCOMStates;
  std::deque<FootAbsolutePosition> m_FIFOLeftFootPosition, m_FIFORightFootPosition;
};

/* patterngeneratorinterface.hh:306 */
PatternGeneratorInterface *patternGeneratorInterfaceFactory(CjrlHumanoidDynamicRobot *aHDR);

}  // namespace PatternGeneratorJRL

/* src/Mathematics/qld.hh:27-31, argument for argument (C++ linkage as in the reference): one dense QP through
 * wg_qld_solve_batch on the process-wide context.  war / iwar are not used (iwar[0] = 0, "C holds its Cholesky factor", is
 * not supported: ifail = 5); eps1 plays QLD's role for the Hessian (vsmall of the diagonal boost, see walkgen_b200.h). */
int ql0001_(int *m, int *me, int *mmax, int *n, int *nmax, int *mnn, double *c, double *d, double *a, double *b, double *xl,
            double *xu, double *x, double *u, int *iout, int *ifail, int *iprint, double *war, int *lwar, int *iwar,
            int *liwar, double *eps1);

#endif
