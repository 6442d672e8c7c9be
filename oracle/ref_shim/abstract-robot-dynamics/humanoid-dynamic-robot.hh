/* Stand-in for <abstract-robot-dynamics/humanoid-dynamic-robot.hh> (abstract-robot-dynamics >= 1.15
 * is an un-vendored dependency of the reference, CMakeLists.txt:46).  Written for this repository:
 * the three foot accessors FootConstraintsAsLinearSystem.cpp:269-281 calls, and the state / property /
 * ZMP accessors ZMPPreviewControlWithMultiBodyZMP.cpp calls (:163-198, :447-479, :487-506, :563-578).
 * The "robot" holds no model: zeroMomentumPoint() returns the entry of a caller-supplied stream selected
 * by the iteration number the test realisation (oracle/ref_glue_twostage.cc) records, which is how a test
 * feeds the reference's second stage a multibody ZMP of its choosing.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_REF_SHIM_ARD_HH
#define ORACLE_REF_SHIM_ARD_HH
#include <string>
#include <jrl/mal/matrixabstractlayer.hh>
struct vector3d {
  double v[3];
  double &operator[](int i) { return v[i]; }
};
typedef oracle_mal::vector<double> vectorN;
class CjrlFoot {
 public:
  double sole_length, sole_width, ankle_z;
  void getSoleSize(double &outLength, double &outWidth) const { outLength = sole_length; outWidth = sole_width; }
  void getAnklePositionInLocalFrame(vector3d &out) const { out[0] = 0; out[1] = 0; out[2] = ankle_z; }
};
class CjrlHumanoidDynamicRobot {
 public:
  CjrlHumanoidDynamicRobot() : zmp_stream(0), zmp_stream_len(0), iteration(0), q(36), dq(36), ddq(36) {}
  CjrlFoot right, left;
  CjrlFoot *rightFoot() { return &right; }
  CjrlFoot *leftFoot() { return &left; }
  /* multibody ZMP stream [len][2] and the stage-1 iteration whose posture was realised last */
  const double *zmp_stream;
  long zmp_stream_len;
  long iteration;
  vectorN q, dq, ddq;
  const vectorN &currentConfiguration() const { return q; }
  const vectorN &currentVelocity() const { return dq; }
  const vectorN &currentAcceleration() const { return ddq; }
  bool currentConfiguration(const vectorN &v) { q = v; return true; }
  bool currentVelocity(const vectorN &v) { dq = v; return true; }
  bool currentAcceleration(const vectorN &v) { ddq = v; return true; }
  bool setProperty(std::string &, const std::string &) { return true; }
  bool getProperty(const std::string &, std::string &) { return true; }
  bool computeForwardKinematics() { return true; }
  oracle_mal::vec3<double> zeroMomentumPoint() const
  {
    oracle_mal::vec3<double> z;
    if (zmp_stream && iteration >= 0 && iteration < zmp_stream_len) {
      z[0] = zmp_stream[2 * iteration]; z[1] = zmp_stream[2 * iteration + 1];
    } else { z[0] = z[1] = 0.0 / 0.0; }
    return z;
  }
  oracle_mal::vec3<double> positionCenterOfMass() const { return oracle_mal::vec3<double>(); }
};
#endif
