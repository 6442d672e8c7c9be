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

#define MAL_MATRIX(name, type) walkgen_b200::Matrix name
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

class ZMPRefTrajectoryGeneration : public SimplePlugin {
 public:
  explicit ZMPRefTrajectoryGeneration(SimplePluginManager *lSPM);
  virtual ~ZMPRefTrajectoryGeneration() {}
  virtual int InitOnLine(std::deque<ZMPPosition> &FinalZMPPositions, std::deque<COMState> &CoMStates,
                         std::deque<FootAbsolutePosition> &FinalLeftFootTraj_deq,
                         std::deque<FootAbsolutePosition> &FinalRightFootTraj_deq,
                         FootAbsolutePosition &InitLeftFootAbsolutePosition,
                         FootAbsolutePosition &InitRightFootAbsolutePosition, std::deque<double> &RelativeFootPositions,
                         COMState &lStartingCOMState, double lStartingZMPPosition[3]) = 0;
  virtual void OnLine(double time, std::deque<ZMPPosition> &FinalZMPPositions, std::deque<COMState> &CoMStates,
                      std::deque<FootAbsolutePosition> &FinalLeftFootTraj_deq,
                      std::deque<FootAbsolutePosition> &FinalRightFootTraj_deq) = 0;
  virtual void CallMethod(std::string &Method, std::istringstream &strm);
  double GetTSingleSupport() const { return m_Tsingle; }
  double GetTDoubleSupport() const { return m_Tdble; }
  double GetSamplingPeriod() const { return m_SamplingPeriod; }
  double GetComHeight() const { return m_ComHeight; }
  bool GetOnLineMode() const { return m_OnLineMode; }
 protected:
  double m_Tsingle, m_Tdble, m_SamplingPeriod, m_Omega, m_ComHeight, m_StepHeight;
  bool m_OnLineMode;
};

class ZMPVelocityReferencedQP : public ZMPRefTrajectoryGeneration {
 public:
  /* the robot is only asked for its sole size in the reference (RelativeFeetInequalities): pass it directly */
  ZMPVelocityReferencedQP(SimplePluginManager *lSPM, std::string DataFile, double sole_length = 0.25,
                          double sole_width = 0.14);
  ~ZMPVelocityReferencedQP();
  int InitOnLine(std::deque<ZMPPosition> &FinalZMPPositions, std::deque<COMState> &CoMStates,
                 std::deque<FootAbsolutePosition> &FinalLeftFootTraj_deq,
                 std::deque<FootAbsolutePosition> &FinalRightFootTraj_deq,
                 FootAbsolutePosition &InitLeftFootAbsolutePosition, FootAbsolutePosition &InitRightFootAbsolutePosition,
                 std::deque<double> &RelativeFootPositions, COMState &lStartingCOMState, double lStartingZMPPosition[3]);
  void OnLine(double time, std::deque<ZMPPosition> &FinalZMPPositions, std::deque<COMState> &CoMStates,
              std::deque<FootAbsolutePosition> &FinalLeftFootTraj_deq,
              std::deque<FootAbsolutePosition> &FinalRightFootTraj_deq);
  void Reference(std::istringstream &strm) { strm >> m_State.new_ref[0] >> m_State.new_ref[1] >> m_State.new_ref[2]; }
  void Reference(double dx, double dy, double dyaw) { m_State.new_ref[0] = dx; m_State.new_ref[1] = dy; m_State.new_ref[2] = dyaw; }
  bool Running() const { return m_State.running != 0; }
  void EndingPhase(bool EndingPhase) { m_State.ending_phase = EndingPhase; }
  void setCoMPerturbationForce(double, double) {}   /* parsed but never consumed by the reference either */
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
};

/* ---- Dimitrov2008 path: the classes around PLDPSolver (names and signatures of the reference) ------------------------ */
struct RelativeFootPosition {            /* include/jrl/walkgen/pgtypes.hh:100-109 */
  double sx, sy, theta;
  double SStime, DStime;
  int stepType;
  double DeviationHipHeight;
};
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
                            COMState &lStartingCOMState, double lStartingZMPPosition[3],
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
                 std::deque<FootAbsolutePosition> &, FootAbsolutePosition &, FootAbsolutePosition &, std::deque<double> &,
                 COMState &, double[3]) { return 0; }
  void OnLine(double, std::deque<ZMPPosition> &, std::deque<COMState> &, std::deque<FootAbsolutePosition> &,
              std::deque<FootAbsolutePosition> &) {}
 private:
  wg_dimitrov_params m_Par;
  wg_zmpdisc_params m_Zd;
  bool m_Dirty;
  int m_Status, m_Done;
};

/* The PGI facade for the Herdt path: command bus + the 5 ms tick. */
class PatternGeneratorInterface : public SimplePluginManager, public SimplePlugin {
 public:
  PatternGeneratorInterface(double sole_length = 0.25, double sole_width = 0.14);
  ~PatternGeneratorInterface();
  /* Reads the first token and broadcasts the rest (PatternGeneratorInterfacePrivate.cpp:1030-1041). */
  int ParseCmd(std::istringstream &strm);
  void CallMethod(std::string &Method, std::istringstream &strm);
  /* One 5 ms tick (PatternGeneratorInterfacePrivate.cpp:1246-1336, Herdt branch + CoMAndFootOnlyStrategy pop).
   * Returns false when the deques ran empty (end of the motion). */
  bool RunOneStepOfTheControlLoop(COMState &COMStateOut, ZMPPosition &ZMPTarget, FootAbsolutePosition &LeftFootPosition,
                                  FootAbsolutePosition &RightFootPosition);
  void setVelocityReference(double x, double y, double yaw) { m_ZMPVRQP->Reference(x, y, yaw); }
  void SetStartConfiguration(const COMState &com, const FootAbsolutePosition &lf, const FootAbsolutePosition &rf);
  ZMPVelocityReferencedQP *VRQP() { return m_ZMPVRQP; }
  PreviewControl *PC() { return m_PC; }
 private:
  int initOnlineHerdt();
  ZMPVelocityReferencedQP *m_ZMPVRQP;
  PreviewControl *m_PC;
  std::deque<ZMPPosition> m_ZMPPositions;
  std::deque<COMState> m_COMBuffer;
  std::deque<FootAbsolutePosition> m_LeftFootPositions, m_RightFootPositions;
  COMState m_StartCOM;
  FootAbsolutePosition m_StartLF, m_StartRF;
  double m_InternalClock;
  int m_AlgorithmforZMPCOM;   /* 0 Kajita (default), 1 Herdt */
  bool m_Running;
};

}  // namespace PatternGeneratorJRL

#endif
