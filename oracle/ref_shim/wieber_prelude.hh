/* Force-included ahead of ZMPQPWithConstraint.cpp / ZMPRefTrajectoryGeneration.cpp (oracle/Makefile).  The Wieber2006
 * generator only uses its ZMPDiscretization member to produce the feet / ZMP-reference buffers (GetZMPDiscretization,
 * ZMPQPWithConstraint.cpp:1355-1364), which the tests supply themselves; everything the test pins -
 * BuildLinearConstraintInequalities, BuildMatricesPxPu, BuildZMPTrajectoryFromFootTrajectory - needs only the MAL macros,
 * ComputeConvexHull, the two CjrlFoot accessors and ql0001_.  Pre-defining the include guards of ZMPDiscretization.hh and
 * StepStackHandler.hh and declaring a do-nothing ZMPDiscretization skips the foot-trajectory / robot-dynamics / step-over
 * headers without touching the reference files.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_REF_SHIM_WIEBER_PRELUDE_HH
#define ORACLE_REF_SHIM_WIEBER_PRELUDE_HH
#define _STEP_STACK_HANDLER_H_
#define _ZMP_DISCRETIZATION_H_
#include <deque>
#include <string>
#include <jrl/mal/matrixabstractlayer.hh>
#include <abstract-robot-dynamics/humanoid-dynamic-robot.hh>
#include <jrl/walkgen/pgtypes.hh>
namespace PatternGeneratorJRL {
using std::deque;
using std::string;
class StepStackHandler;
class SimplePluginManager;
class ZMPDiscretization {
 public:
  ZMPDiscretization(SimplePluginManager *, std::string, CjrlHumanoidDynamicRobot *) {}
  void GetZMPDiscretization(deque<ZMPPosition> &, deque<COMState> &, deque<RelativeFootPosition> &,
                            deque<FootAbsolutePosition> &, deque<FootAbsolutePosition> &, double, COMState &,
                            MAL_S3_VECTOR(&, double), FootAbsolutePosition &, FootAbsolutePosition &) {}
};
}
#endif
