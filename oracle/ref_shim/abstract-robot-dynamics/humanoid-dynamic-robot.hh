/* Stand-in for <abstract-robot-dynamics/humanoid-dynamic-robot.hh> (abstract-robot-dynamics >= 1.15
 * is an un-vendored dependency of the reference, CMakeLists.txt:46).  Written for this repository:
 * only the three accessors FootConstraintsAsLinearSystem.cpp:269-281 calls.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_REF_SHIM_ARD_HH
#define ORACLE_REF_SHIM_ARD_HH
struct vector3d {
  double v[3];
  double &operator[](int i) { return v[i]; }
};
class CjrlFoot {
 public:
  double sole_length, sole_width, ankle_z;
  void getSoleSize(double &outLength, double &outWidth) const { outLength = sole_length; outWidth = sole_width; }
  void getAnklePositionInLocalFrame(vector3d &out) const { out[0] = 0; out[1] = 0; out[2] = ankle_z; }
};
class CjrlHumanoidDynamicRobot {
 public:
  CjrlFoot right, left;
  CjrlFoot *rightFoot() { return &right; }
  CjrlFoot *leftFoot() { return &left; }
};
#endif
