/* Stand-in for jrl-mal's <jrl/mal/matrixabstractlayer.hh> (jrl-mal >= 1.9.0 is an
 * un-vendored dependency of the reference, CMakeLists.txt:45).  The reference math
 * sources compiled by oracle/Makefile (OptCholesky.cpp, PLDPSolver.cpp) include this
 * header but use none of its macros, so an empty header is sufficient.
 * TEST INFRASTRUCTURE ONLY - nothing under oracle/ is linked into the product. */
#ifndef ORACLE_REF_SHIM_MAL_HH
#define ORACLE_REF_SHIM_MAL_HH
#endif
