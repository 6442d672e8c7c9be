/* Force-included ahead of ZMPConstrainedQPFastFormulation.cpp (oracle/Makefile).  As wieber_prelude.hh: the generator's
 * ZMPDiscretization member only produces the feet / ZMP-reference buffers, which the tests supply; ZMPDiscretization.hh is where
 * the reference picks up the declaration of PreviewControl (a pointer member the generator never dereferences, :261).
 * TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_REF_SHIM_DIMITROV_PRELUDE_HH
#define ORACLE_REF_SHIM_DIMITROV_PRELUDE_HH
#include "wieber_prelude.hh"
namespace PatternGeneratorJRL {
class PreviewControl;
}
#endif
