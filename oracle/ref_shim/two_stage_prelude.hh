/* Force-included ahead of ZMPPreviewControlWithMultiBodyZMP.cpp (oracle/Makefile): ComAndFootRealization.hh:35 includes
 * StepStackHandler.hh, which drags in the step-over planner, the robot dynamics and the foot-trajectory classes, although
 * ComAndFootRealization only stores a StepStackHandler POINTER.  Pre-defining that header's include guard and declaring the
 * class skips it without touching the reference files.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_REF_SHIM_TWO_STAGE_PRELUDE_HH
#define ORACLE_REF_SHIM_TWO_STAGE_PRELUDE_HH
#define _STEP_STACK_HANDLER_H_
#define _ZMP_DISCRETIZATION_H_
#include <deque>
#include <jrl/walkgen/pgtypes.hh>
namespace PatternGeneratorJRL {
class StepStackHandler;
class PatternGeneratorInterfacePrivate;
using std::deque;
}
#endif
