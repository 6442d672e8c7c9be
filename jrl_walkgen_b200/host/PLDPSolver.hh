// PLDPSolver.hh - same header name as the reference; the declarations live in walkgen_host.hh
#include "walkgen_host.hh"
