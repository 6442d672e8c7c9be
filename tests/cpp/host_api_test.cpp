// host_api_test.cpp - drives the host-side C++ class mirror (jrl_walkgen_b200/host) the way the reference's own tests
// drive jrl-walkgen: tests/TestOptCholesky.cpp, tests/TestHerdt2010.cpp + tests/TestObject.cpp, tests/TestRiccatiEquation.cpp.
// Run by tests/test_host_cpp_gpu.py on the GPU box; numerical comparison against the golden datref / the oracle is
// done there on the files this program writes.
//   host_api_test optcholesky
//   host_api_test herdt2010 <out.dat> <nticks> [emergency]
//   host_api_test preview <out.bin>
//   host_api_test pldp <in.bin> <out.bin>
//   host_api_test preview1d <gains.ini> <in.bin> <out.bin>
//   host_api_test kajita2003 <profile> <out.dat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include "../../jrl_walkgen_b200/host/PreviewControl.hh"
#include "../../jrl_walkgen_b200/host/OptCholesky.hh"
#include "../../jrl_walkgen_b200/host/PLDPSolver.hh"
#include "../../jrl_walkgen_b200/host/patterngeneratorinterface.hh"

using namespace PatternGeneratorJRL;

// ---- tests/TestOptCholesky.cpp:86-180 ---------------------------------------------------------------------
static int test_optcholesky()
{
  const unsigned NbOfConstraints = 12, CardU = 15;
  srand(0);
  std::vector<double> A(NbOfConstraints * CardU);
  for (unsigned i = 0; i < NbOfConstraints * CardU; ++i) A[i] = (double)rand() / (double)RAND_MAX;
  std::vector<double> L(NbOfConstraints * NbOfConstraints, 0.0), iL(NbOfConstraints * NbOfConstraints, 0.0);
  OptCholesky anOCD(NbOfConstraints, CardU, OptCholesky::MODE_NORMAL);
  anOCD.SetA(A.data(), NbOfConstraints);
  anOCD.SetL(L.data());
  anOCD.SetiL(iL.data());
  for (unsigned i = 0; i < NbOfConstraints; ++i) anOCD.AddActiveConstraint(i);
  if (anOCD.CurrentNumberOfRows() != (int)NbOfConstraints) return 1;
  double dist = 0.0;
  for (unsigned i = 0; i < NbOfConstraints; ++i)
    for (unsigned j = 0; j < NbOfConstraints; ++j) {
      double m = 0.0, l = 0.0;
      for (unsigned k = 0; k < CardU; ++k) m += A[i * CardU + k] * A[j * CardU + k];
      for (unsigned k = 0; k < NbOfConstraints; ++k) l += L[i * NbOfConstraints + k] * L[j * NbOfConstraints + k];
      dist += (m - l) * (m - l);
    }
  dist = sqrt(dist);
  std::cout << "distance A A^T - L L^T: " << dist << std::endl;
  if (dist > 1e-6) return 2;                                       // TestOptCholesky.cpp:143-150
  // full decomposition of M = A A^T and its inverse (:155-177)
  const unsigned n = NbOfConstraints;
  std::vector<double> M(n * n), L2(n * n, 0.0), iL2(n * n, 0.0);
  for (unsigned i = 0; i < n; ++i)
    for (unsigned j = 0; j < n; ++j) {
      double m = 0.0;
      for (unsigned k = 0; k < CardU; ++k) m += A[i * CardU + k] * A[j * CardU + k];
      M[i * n + j] = m;
    }
  OptCholesky anOCD2(n, n, OptCholesky::MODE_NORMAL);
  anOCD2.SetA(M.data(), n);
  anOCD2.SetL(L2.data());
  anOCD2.SetiL(iL2.data());
  if (anOCD2.ComputeNormalCholeskyOnANormal() != 0) return 3;
  if (anOCD2.ComputeInverseCholeskyNormal(1) != 0) return 4;
  double worst = 0.0;
  for (unsigned i = 0; i < n; ++i)
    for (unsigned j = 0; j < n; ++j) {
      double s = 0.0;
      for (unsigned k = 0; k < n; ++k) s += iL2[i * n + k] * L2[k * n + j];
      worst = std::max(worst, fabs(s - (i == j ? 1.0 : 0.0)));
    }
  std::cout << "max |iL L - I|: " << worst << std::endl;
  return worst < 1e-9 ? 0 : 5;
}

// ---- tests/TestHerdt2010.cpp (OnLine profile) through ParseCmd + the 5 ms tick ------------------------------
static void cmd(PatternGeneratorInterface &pgi, const char *s)
{
  std::istringstream strm(s);
  pgi.ParseCmd(strm);
}
static int test_herdt2010(const char *out, int nticks, bool emergency)
{
  PatternGeneratorInterface aPGI;
  // tests/CommonTools.cpp:56-76 (the first 9 commands, as TestHerdt2010 sends them)
  const char *common[] = {":comheight 0.8078", ":samplingperiod 0.005", ":previewcontroltime 1.6", ":omega 0.0",
                          ":stepheight 0.07", ":singlesupporttime 0.78", ":doublesupporttime 0.02", ":armparameters 0.5",
                          ":LimitsFeasibility 0.0"};
  for (const char *c : common) cmd(aPGI, c);
  // tests/TestHerdt2010.cpp:66-90
  cmd(aPGI, ":SetAlgoForZmpTrajectory Herdt");
  cmd(aPGI, ":singlesupporttime 0.7");
  cmd(aPGI, ":doublesupporttime 0.1");
  cmd(aPGI, emergency ? ":HerdtOnline 0.2 0.0 0.2" : ":HerdtOnline 0.2 0.0 0.0");
  cmd(aPGI, ":numberstepsbeforestop 2");
  aPGI.VRQP()->SetDatrefEra();                                    // the code/robot the datref was made with
  std::ofstream aof(out);
  aof.precision(8);
  aof.setf(std::ios::scientific, std::ios::floatfield);
  COMState com; ZMPPosition zmp; FootAbsolutePosition lf, rf;
  int it = 0;
  for (; it < nticks; ++it) {
    if (!aPGI.RunOneStepOfTheControlLoop(com, zmp, lf, rf)) break;
    // the 38 columns of tests/TestObject.cpp:333-389
    aof << (it + 1) * 0.005 << " " << com.x[0] << " " << com.y[0] << " " << com.z[0] << " " << com.yaw[0] << " " << com.x[1]
        << " " << com.y[1] << " " << com.z[1] << " " << zmp.px << " " << zmp.py << " ";
    const FootAbsolutePosition *ff[2] = {&lf, &rf};
    for (int f = 0; f < 2; ++f)
      aof << ff[f]->x << " " << ff[f]->y << " " << ff[f]->z << " " << ff[f]->dx << " " << ff[f]->dy << " " << ff[f]->dz << " "
          << ff[f]->ddx << " " << ff[f]->ddy << " " << ff[f]->ddz << " " << ff[f]->theta << " " << ff[f]->omega << " "
          << ff[f]->omega2 << " ";
    aof << zmp.px << " " << zmp.py << " 0 0" << std::endl;
    // generateEvent(), tests/TestHerdt2010.cpp:232-244
    if (emergency) {   // generateEventEmergencyStop(), :258-262
      if (it == 5 * 200) cmd(aPGI, ":setVelReference  0.0 0.0 0.4");
      if (it == 10 * 200) cmd(aPGI, ":setVelReference  0.2 0.0 -0.2");
      if (it == 3040) cmd(aPGI, ":setVelReference 0.0 0.0 0.0");
      if (it == 4160) { cmd(aPGI, ":setVelReference  0.0 0.0 0.0"); cmd(aPGI, ":stoppg"); }
      continue;
    }
    if (it == 5 * 200 || it == 35 * 200 || it == 55 * 200 || it == 75 * 200) cmd(aPGI, ":setVelReference  0.2 0.0 0.0");
    if (it == 10 * 200) cmd(aPGI, ":setVelReference  0.0 0.2 0.0");
    if (it == 25 * 200 || it == 65 * 200) cmd(aPGI, ":setVelReference  0.0 0.0 -10.");
    if (it == 45 * 200) cmd(aPGI, ":setVelReference  0.0 0.0 10.0");
    if (it == 85 * 200) cmd(aPGI, ":setVelReference  0.2 0.0 6.0832");
    if (it == 95 * 200) cmd(aPGI, ":setVelReference  0.2 0.0 -6.0832");
    if (it == 105 * 200) cmd(aPGI, ":setVelReference 0.0 0.0 0.0");
    if (it == 110 * 200) { cmd(aPGI, ":setVelReference  0.0 0.0 0.0"); cmd(aPGI, ":stoppg"); }
  }
  std::cout << "ticks written: " << it << std::endl;
  return it == nticks ? 0 : 1;
}

// ---- PreviewControl: tick-by-tick OneIterationOfPreview vs the batched whole-trajectory call -----------------
static int test_preview(const char *out)
{
  SimplePluginManager spm;
  PreviewControl aPC(&spm, OptimalControllerSolver::MODE_WITHOUT_INITIALPOS, true);
  std::string m1(":samplingperiod"), m2(":previewcontroltime"), m3(":comheight");
  { std::istringstream s("0.005"); spm.CallMethod(m1, s); }
  { std::istringstream s("1.6"); spm.CallMethod(m2, s); }
  { std::istringstream s("0.814"); spm.CallMethod(m3, s); }
  if (!aPC.IsCoherent()) return 1;
  std::cout << "Ks " << aPC.Gains().Ks << " Kx " << aPC.Gains().Kx[0] << " " << aPC.Gains().Kx[1] << " " << aPC.Gains().Kx[2]
            << " F0 " << aPC.Gains().F[0] << std::endl;
  const unsigned NL = 320, nsteps = 200;
  std::deque<ZMPPosition> zmp(NL + nsteps);
  for (unsigned i = 0; i < zmp.size(); ++i) {
    zmp[i].px = 0.2 * (i / 160); zmp[i].py = ((i / 160) % 2 ? -0.095 : 0.095); zmp[i].pz = 0; zmp[i].theta = 0;
    zmp[i].time = i * 0.005; zmp[i].stepType = 1;
  }
  MAL_MATRIX_DIM(x, double, 3, 1); MAL_MATRIX_DIM(y, double, 3, 1);
  double sx = 0, sy = 0, zx = 0, zy = 0;
  std::vector<double> serial;
  for (unsigned k = 0; k < nsteps; ++k) {
    aPC.OneIterationOfPreview(x, y, sx, sy, zmp, k, zx, zy, true);
    for (int i = 0; i < 3; ++i) serial.push_back(x(i, 0));
    for (int i = 0; i < 3; ++i) serial.push_back(y(i, 0));
    serial.push_back(zx); serial.push_back(zy);
  }
  MAL_MATRIX_DIM(x2, double, 3, 1); MAL_MATRIX_DIM(y2, double, 3, 1);
  double sx2 = 0, sy2 = 0;
  std::vector<double> com6, zmp2;
  std::deque<ZMPPosition> zcut(zmp.begin(), zmp.begin() + NL + nsteps - 1);
  const int steps = aPC.RunWholeTrajectory(zcut, x2, y2, sx2, sy2, com6, zmp2, true);
  if (steps != (int)nsteps) return 2;
  double worst = 0.0;
  for (unsigned k = 0; k < nsteps; ++k) {
    for (int i = 0; i < 6; ++i) worst = std::max(worst, fabs(com6[6 * k + i] - serial[8 * k + i]));
    worst = std::max(worst, fabs(zmp2[2 * k] - serial[8 * k + 6]));
  }
  std::cout << "max |batched - tick by tick|: " << worst << std::endl;
  // window under-filled: the reference LTHROWs (PreviewControl.cpp:341-344)
  bool thrown = false;
  std::deque<ZMPPosition> shortq(zmp.begin(), zmp.begin() + NL - 1);
  try { aPC.OneIterationOfPreview(x, y, sx, sy, shortq, 0, zx, zy, true); } catch (const std::exception &) { thrown = true; }
  if (!thrown) return 3;
  // 1-D variant on a deque<double>
  std::deque<double> z1(NL + 5);
  for (unsigned i = 0; i < z1.size(); ++i) z1[i] = zmp[i].px;
  MAL_MATRIX_DIM(x1, double, 3, 1);
  double s1 = 0, zz = 0;
  aPC.OneIterationOfPreview1D(x1, s1, z1, 0, zz, true);
  if (fabs(x1(0, 0) - serial[0]) > 1e-15) return 4;
  std::ofstream f(out, std::ios::binary);
  f.write(reinterpret_cast<const char *>(serial.data()), sizeof(double) * serial.size());
  return worst < 1e-10 ? 0 : 5;
}

// ---- tests/TestKajita2003.cpp through ParseCmd + the 5 ms tick ------------------------------------------------------
// profile: StraightWalking (:stepseq, tests/TestKajita2003.cpp:104-119), Circle (:supportfoot / :arc / :lastsupport /
// :finish, :68-92), OnLine (:StartOnLineStepSequencing + :StopOnLineStepSequencing after 5 s).  Writes one row per tick:
// time, CoM x y z yaw dx dy (7), ZMP ref px py (2), left foot x y z theta omega omega2 (6), right foot (6) = 22 columns.
static int test_kajita2003(const char *profile, const char *out)
{
  CjrlHumanoidDynamicRobot robot;                                  // sole 0.25 x 0.14, ankle height 0.105
  PatternGeneratorInterface *aPGI = patternGeneratorInterfaceFactory(&robot);
  // tests/CommonTools.cpp:56-76
  const char *common[] = {":comheight 0.8078", ":samplingperiod 0.005", ":previewcontroltime 1.6", ":omega 0.0",
                          ":stepheight 0.07", ":singlesupporttime 0.78", ":doublesupporttime 0.02", ":armparameters 0.5",
                          ":LimitsFeasibility 0.0", ":ZMPShiftParameters 0.015 0.015 0.015 0.015",
                          ":TimeDistributionParameters 2.0 3.7 1.7 3.0", ":UpperBodyMotionParameters -0.1 -1.0 0.0"};
  for (const char *c : common) cmd(*aPGI, c);
  // start configuration of the HRP-2 half-sitting pose as the reference evaluates it (first row of the Kajita datrefs)
  COMState com0; FootAbsolutePosition lf0, rf0;
  std::memset(&lf0, 0, sizeof lf0); std::memset(&rf0, 0, sizeof rf0);
  com0.z[0] = 0.8078; lf0.x = rf0.x = 0.00949035; lf0.y = 0.095; rf0.y = -0.095;
  aPGI->SetStartConfiguration(com0, lf0, rf0);
  const std::string prof(profile);
  cmd(*aPGI, ":SetAlgoForZmpTrajectory Kajita");
  if (prof == "StraightWalking") {
    cmd(*aPGI, ":stepseq 0.0 -0.105 0.0 0.2 0.21 0.0 0.2 -0.21 0.0 0.2 0.21 0.0 0.2 -0.21 0.0 0.2 0.21 0.0 0.2 -0.21 0.0 "
               "0.2 0.21 0.0 0.2 -0.21 0.0 0.2 0.21 0.0 0.2 -0.21 0.0 0.2 0.21 0.0 0.2 -0.21 0.0 0.2 0.21 0.0 0.2 -0.21 0.0 0.0 0.21 0.0");
  } else if (prof == "Circle") {
    cmd(*aPGI, ":supportfoot 1");
    cmd(*aPGI, ":arc 0.0 0.75 30.0 -1");
    cmd(*aPGI, ":lastsupport");
    cmd(*aPGI, ":finish");
  } else if (prof == "OnLine") {
    cmd(*aPGI, ":StartOnLineStepSequencing 0.0 -0.105 0.0 0.2 0.21 0.0 0.2 -0.21 0.0 0.2 0.21 0.0");
  } else {
    return 2;
  }
  std::ofstream aof(out);
  aof.precision(10);
  aof.setf(std::ios::scientific, std::ios::floatfield);
  // the vector overload of patterngeneratorinterface.hh:152-158, as TestObject.cpp:280-300 calls it
  std::vector<double> conf(36, 0.0), vel(36, 0.0), acc(36, 0.0), zmptarget(3, 0.0);
  COMState com; FootAbsolutePosition lf, rf;
  int it = 0;
  for (; it < 40000; ++it) {
    if (prof == "OnLine" && it == 1000) cmd(*aPGI, ":StopOnLineStepSequencing");
    if (!aPGI->RunOneStepOfTheControlLoop(conf, vel, acc, zmptarget, com, lf, rf)) break;
    aof << (it + 1) * 0.005 << " " << com.x[0] << " " << com.y[0] << " " << com.z[0] << " " << com.yaw[0] << " " << com.x[1]
        << " " << com.y[1] << " " << zmptarget[0] << " " << zmptarget[1];
    const FootAbsolutePosition *ff[2] = {&lf, &rf};
    for (int f = 0; f < 2; ++f)
      aof << " " << ff[f]->x << " " << ff[f]->y << " " << ff[f]->z << " " << ff[f]->theta << " " << ff[f]->omega << " "
          << ff[f]->omega2;
    aof << std::endl;
  }
  for (double v : conf) if (v != 0.0) return 3;                   // joint space is left untouched on this path
  std::cout << "ticks written: " << it << std::endl;
  delete aPGI;
  return it > 0 ? 0 : 1;
}

// ---- PreviewControl: ReadPrecomputedFile + both OneIterationOfPreview1D overloads on inputs from a file --------------
// in.bin : int32 L, ncases; double z[L] (deque overload, run with lindex = k over the whole buffer from a zero state);
//          then per case int32 Lc, lindex, sim, pad; double x0[3], s0, buf[Lc] (vector overload, one call).
// out.bin: per deque step x[3], zmp (4 doubles); per case x[3], s, zmp (5 doubles); then one double: mean wall-clock
//          microseconds of a OneIterationOfPreview call (2-D, per-tick path of the class mirror).
static int test_preview1d(const char *ini, const char *in, const char *out)
{
  SimplePluginManager spm;
  PreviewControl aPC(&spm, OptimalControllerSolver::MODE_WITHOUT_INITIALPOS, false);
  aPC.ReadPrecomputedFile(ini);
  if (!aPC.IsCoherent()) return 1;
  const unsigned NL = (unsigned)aPC.Gains().NL;
  std::ifstream f(in, std::ios::binary);
  int32_t hdr[2];
  f.read(reinterpret_cast<char *>(hdr), sizeof hdr);
  std::vector<double> z(hdr[0]);
  f.read(reinterpret_cast<char *>(z.data()), 8 * z.size());
  std::ofstream o(out, std::ios::binary);
  std::deque<double> zq(z.begin(), z.end());
  MAL_MATRIX_DIM(x, double, 3, 1);
  double s = 0.0, zz = 0.0;
  for (unsigned k = 0; k + NL <= zq.size(); ++k) {
    aPC.OneIterationOfPreview1D(x, s, zq, k, zz, true);
    const double row[4] = {x(0, 0), x(1, 0), x(2, 0), zz};
    o.write(reinterpret_cast<const char *>(row), sizeof row);
  }
  for (int c = 0; c < hdr[1]; ++c) {
    int32_t h4[4];
    f.read(reinterpret_cast<char *>(h4), sizeof h4);
    double x0s[4];
    f.read(reinterpret_cast<char *>(x0s), sizeof x0s);
    std::vector<double> buf(h4[0]);
    f.read(reinterpret_cast<char *>(buf.data()), 8 * buf.size());
    MAL_MATRIX_DIM(xc, double, 3, 1);
    for (int i = 0; i < 3; ++i) xc(i, 0) = x0s[i];
    double sc = x0s[3], zc = 0.0;
    aPC.OneIterationOfPreview1D(xc, sc, buf, (unsigned)h4[1], zc, h4[2] != 0);
    const double row[5] = {xc(0, 0), xc(1, 0), xc(2, 0), sc, zc};
    o.write(reinterpret_cast<const char *>(row), sizeof row);
  }
  // latency of the per-tick path (what a PGI calling the class mirror every 5 ms pays)
  std::deque<ZMPPosition> zmp(NL + 8);
  for (unsigned i = 0; i < zmp.size(); ++i) { zmp[i].px = 0.01 * i; zmp[i].py = 0.1; zmp[i].pz = 0; zmp[i].theta = 0; zmp[i].time = 0; zmp[i].stepType = 1; }
  MAL_MATRIX_DIM(x2, double, 3, 1); MAL_MATRIX_DIM(y2, double, 3, 1);
  double sx = 0, sy = 0, zx = 0, zy = 0;
  for (int k = 0; k < 50; ++k) aPC.OneIterationOfPreview(x2, y2, sx, sy, zmp, 0, zx, zy, true);
  const int reps = 2000;
  const auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < reps; ++k) aPC.OneIterationOfPreview(x2, y2, sx, sy, zmp, k & 7, zx, zy, true);
  const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
  o.write(reinterpret_cast<const char *>(&us), sizeof us);
  std::cout << "OneIterationOfPreview per tick: " << us << " us" << std::endl;
  return 0;
}

// ---- PLDPSolver: one problem from a file, through the reference's class interface ---------------------------
static int test_pldp(const char *in, const char *out)
{
  std::ifstream f(in, std::ios::binary);
  int32_t hdr[2];
  f.read(reinterpret_cast<char *>(hdr), sizeof hdr);
  const unsigned N = 16, nprob = (unsigned)hdr[1];
  std::vector<double> iPu(N * N), Px(N * 3), Pu(N * N), iLQ(4 * N * N, 0.0);
  f.read(reinterpret_cast<char *>(iPu.data()), 8 * iPu.size());
  f.read(reinterpret_cast<char *>(Px.data()), 8 * Px.size());
  f.read(reinterpret_cast<char *>(Pu.data()), 8 * Pu.size());
  Optimization::Solver::PLDPSolver solver(N, iPu.data(), Px.data(), Pu.data(), iLQ.data());
  std::ofstream o(out, std::ios::binary);
  std::vector<int> similar(8 * N, 0);
  for (unsigned p = 0; p < nprob; ++p) {
    int32_t mm = 0, nrem = 0, start = 0, pad = 0;
    f.read(reinterpret_cast<char *>(&mm), 4); f.read(reinterpret_cast<char *>(&nrem), 4);
    f.read(reinterpret_cast<char *>(&start), 4); f.read(reinterpret_cast<char *>(&pad), 4);
    const unsigned m = (unsigned)mm;
    std::vector<double> D(2 * N), DPu((m + 1) * 2 * N), DPx(m), Z(2 * N), xk(6), X(2 * N);
    f.read(reinterpret_cast<char *>(D.data()), 8 * D.size());
    f.read(reinterpret_cast<char *>(DPu.data()), 8 * DPu.size());
    f.read(reinterpret_cast<char *>(DPx.data()), 8 * DPx.size());
    f.read(reinterpret_cast<char *>(Z.data()), 8 * Z.size());
    f.read(reinterpret_cast<char *>(xk.data()), 8 * xk.size());
    const int rc = solver.SolveProblem(D.data(), m, DPu.data(), DPx.data(), Z.data(), xk.data(), X.data(), similar,
                                       (unsigned)nrem, start != 0);
    if (rc != 0) return 10 + p;
    o.write(reinterpret_cast<const char *>(X.data()), 8 * X.size());
  }
  return 0;
}


// ZMPConstrainedQPFastFormulation::GetZMPDiscretization on TestKajita2003's straight walk (tests/TestKajita2003.cpp:104-119)
// + FootConstraintsAsLinearSystem on the feet it returns.  out: int64 n, int32 status, done, npoly, pad; com[n][6];
// rows[npoly] (int32)
static int test_dimitrov(const char *out, bool robust)
{
  SimplePluginManager spm;
  ZMPConstrainedQPFastFormulation gen(&spm, "");
  const char *cmds[] = {":samplingperiod 0.005", ":singlesupporttime 0.78", ":doublesupporttime 0.02", ":stepheight 0.07",
                        ":setdimitrovconstraint 0.04 0.04"};
  for (const char *c : cmds) {
    std::istringstream is(c);
    std::string m; is >> m;
    spm.CallMethod(m, is);
  }
  gen.SetRobustMode(robust, robust);
  std::deque<RelativeFootPosition> rel;
  const double seq[][3] = {{0.0, -0.105, 0.0}, {0.2, 0.21, 0.0}, {0.2, -0.21, 0.0}, {0.2, 0.21, 0.0}, {0.2, -0.21, 0.0},
                           {0.2, 0.21, 0.0}, {0.2, -0.21, 0.0}, {0.2, 0.21, 0.0}, {0.2, -0.21, 0.0}, {0.2, 0.21, 0.0},
                           {0.2, -0.21, 0.0}, {0.2, 0.21, 0.0}, {0.2, -0.21, 0.0}, {0.2, 0.21, 0.0}, {0.2, -0.21, 0.0},
                           {0.0, 0.21, 0.0}};
  for (const auto &t : seq) { RelativeFootPosition r = {t[0], t[1], t[2], 0.78, 0.02, 1, 0.0}; rel.push_back(r); }
  std::deque<ZMPPosition> zmp; std::deque<COMState> com; std::deque<FootAbsolutePosition> left, right;
  COMState start; S3Vector zstart;
  FootAbsolutePosition il, ir;
  std::memset(&il, 0, sizeof il); std::memset(&ir, 0, sizeof ir);
  il.x = 0.00949035; il.y = 0.095; ir.x = 0.00949035; ir.y = -0.095;
  gen.GetZMPDiscretization(zmp, com, rel, left, right, 0.0, start, zstart, il, ir);
  FootConstraintsAsLinearSystem fcals(&spm);
  std::deque<LinearConstraintInequality_t *> q;
  // the step type of the samples is what ZMPDiscretization writes; GetZMPDiscretization does not return it, the polygon
  // scan only needs ">= 10 means double support": mark the samples where both feet are on the ground
  for (size_t i = 0; i < left.size(); ++i) left[i].stepType = (left[i].z == 0.0 && right[i].z == 0.0) ? 10 : 1;
  if (fcals.BuildLinearConstraintInequalities(left, right, q, 0.04, 0.04) != 0) return 3;
  FILE *f = fopen(out, "wb");
  if (!f) return 4;
  const int64_t n = (int64_t)com.size();
  const int32_t hdr[4] = {gen.LastStatus(), gen.PeriodsDone(), (int32_t)q.size(), 0};
  fwrite(&n, 8, 1, f); fwrite(hdr, 4, 4, f);
  for (int64_t i = 0; i < n; ++i) { fwrite(com[i].x, 8, 3, f); fwrite(com[i].y, 8, 3, f); }
  for (size_t k = 0; k < q.size(); ++k) { const int32_t r = (int32_t)q[k]->A.size1(); fwrite(&r, 4, 1, f); delete q[k]; }
  fclose(f);
  std::vector<CH_Point> pts = {{0.1, 0.0}, {0.0, 0.1}, {-0.1, 0.0}, {0.0, -0.1}, {0.0, 0.0}}, hull;
  ComputeConvexHull ch;
  ch.DoComputeConvexHull(pts, hull);
  return hull.size() == 4 ? 0 : 5;
}

// ---- ZMPPreviewControlWithMultiBodyZMP: Setup + OneGlobalStepOfControl tick by tick, as the reference's callers drive it
// (PatternGeneratorInterfacePrivate.cpp, DoubleStagePreviewControlStrategy), with test doubles for the model: the
// realisation records the first-stage CoM it is asked to realise, the robot answers zeroMomentumPoint() with the
// cart-table ZMP of that CoM at 0.9 zc plus a slow disturbance (tests/test_two_stage.py: synthetic_multibody_zmp).
// Input: [L][2] doubles (ZMP reference); output: steps x 6 final CoM, then the batched RunWholeTrajectory result.
namespace {
struct ModelRobot : public CjrlHumanoidDynamicRobot {
  double com6[6]; long tick; double zc;
  ModelRobot() : tick(0), zc(0.807709) { for (int i = 0; i < 6; ++i) com6[i] = 0; }
  static void model(long k, const double *c, double zc, double *out)
  {
    const double h = 0.9 * zc / 9.81;
    out[0] = c[0] - h * c[2] + 0.003 * sin(0.02 * k);
    out[1] = c[3] - h * c[5] + 0.002 * cos(0.015 * k);
  }
  vector3d zeroMomentumPoint() const { vector3d z; double o[2]; model(tick, com6, zc, o); z[0] = o[0]; z[1] = o[1]; z[2] = 0; return z; }
};
struct RecordingRealization : public ComAndFootRealization {
  ModelRobot *robot; double start[3];
  bool ComputePostureForGivenCoMAndFeetPosture(std::vector<double> &p, std::vector<double> &v, std::vector<double> &a,
                                               std::vector<double> &, std::vector<double> &, std::vector<double> &,
                                               std::vector<double> &, std::vector<double> &, int it, int stage)
  {
    if (stage == 0) {
      robot->tick = it;
      robot->com6[0] = p[0]; robot->com6[1] = v[0]; robot->com6[2] = a[0];
      robot->com6[3] = p[1]; robot->com6[4] = v[1]; robot->com6[5] = a[1];
    }
    return true;
  }
  bool InitializationCoM(std::vector<double> &, S3Vector &c, std::vector<double> &, FootAbsolutePosition &, FootAbsolutePosition &)
  { c[0] = start[0]; c[1] = start[1]; c[2] = start[2]; return true; }
};
void model_cb(void *user, long k, const double *c, double *out) { ModelRobot::model(k, c, *static_cast<double *>(user), out); }
}  // namespace

static int test_twostage(const char *in, const char *out, int max_ticks)
{
  std::ifstream fi(in, std::ios::binary);
  std::vector<double> zin((std::istreambuf_iterator<char>(fi)), std::istreambuf_iterator<char>());
  fi.clear(); fi.seekg(0, std::ios::end);
  const size_t bytes = (size_t)fi.tellg();
  fi.seekg(0);
  zin.assign(bytes / sizeof(double), 0.0);
  fi.read(reinterpret_cast<char *>(zin.data()), bytes);
  const size_t L = zin.size() / 2;
  SimplePluginManager spm;
  ZMPPreviewControlWithMultiBodyZMP zpc(&spm);
  PreviewControl pc(&spm, OptimalControllerSolver::MODE_WITHOUT_INITIALPOS, true);
  pc.SetSamplingPeriod(0.005); pc.SetPreviewControlTime(1.6); pc.SetHeightOfCoM(0.807709);
  zpc.SetPreviewControl(&pc);
  ModelRobot robot;
  RecordingRealization cfr; cfr.robot = &robot;
  cfr.start[0] = zin[0] + 0.001; cfr.start[1] = zin[1] - 0.002; cfr.start[2] = 0.807709;
  cfr.setHumanoidDynamicRobot(&robot);
  zpc.setComAndFootRealization(&cfr);
  zpc.setHumanoidDynamicRobot(&robot);
  const unsigned NL = 320;
  std::deque<ZMPPosition> ref(L);
  std::deque<COMState> coms(L);
  std::deque<FootAbsolutePosition> lf(L), rf(L);
  for (size_t i = 0; i < L; ++i) {
    std::memset(&ref[i], 0, sizeof(ZMPPosition)); std::memset(&lf[i], 0, sizeof(FootAbsolutePosition));
    std::memset(&rf[i], 0, sizeof(FootAbsolutePosition));
    ref[i].px = zin[2 * i]; ref[i].py = zin[2 * i + 1]; ref[i].time = 0.005 * i;
  }
  std::vector<double> body(36), waist(6);
  S3Vector sc;
  zpc.EvaluateStartingCoM(body, sc, waist, lf[0], rf[0]);
  zpc.Setup(ref, coms, lf, rf);
  std::vector<double> q, dq, ddq, serial;
  size_t next_ref = 2 * NL + 1;
  int steps = 0;
  while (steps < max_ticks && next_ref <= L) {
    COMState fin; ZMPPosition zp; std::memset(&zp, 0, sizeof zp);
    zpc.OneGlobalStepOfControl(lf[0], rf[0], zp, fin, q, dq, ddq);
    for (int j = 0; j < 3; ++j) serial.push_back(fin.x[j]);
    for (int j = 0; j < 3; ++j) serial.push_back(fin.y[j]);
    ++steps;
    if (next_ref < L) zpc.UpdateTheZMPRefQueue(ref[next_ref]);
    ++next_ref;
  }
  // batched form over the whole stream
  ZMPPreviewControlWithMultiBodyZMP zb(&spm);
  zb.SetPreviewControl(&pc);
  COMState start; start.x[0] = cfr.start[0]; start.y[0] = cfr.start[1]; start.z[0] = cfr.start[2];
  std::deque<COMState> fin_b;
  double zc = 0.807709;
  const int nb = zb.RunWholeTrajectory(ref, start, model_cb, &zc, fin_b);
  if (nb != (int)(L - 2 * NL)) { std::cerr << "batched steps " << nb << std::endl; return 2; }
  double worst = 0.0;
  for (int n = 0; n < steps; ++n)
    for (int j = 0; j < 3; ++j) {
      worst = std::max(worst, fabs(fin_b[n].x[j] - serial[6 * n + j]));
      worst = std::max(worst, fabs(fin_b[n].y[j] - serial[6 * n + 3 + j]));
    }
  std::cout << "two-stage: " << steps << " ticks, max |batched - tick by tick| = " << worst << std::endl;
  std::ofstream f(out, std::ios::binary);
  const double hdr[2] = {(double)steps, (double)nb};
  f.write(reinterpret_cast<const char *>(hdr), sizeof hdr);
  f.write(reinterpret_cast<const char *>(serial.data()), sizeof(double) * serial.size());
  for (int n = 0; n < nb; ++n) {
    double r[6] = {fin_b[n].x[0], fin_b[n].x[1], fin_b[n].x[2], fin_b[n].y[0], fin_b[n].y[1], fin_b[n].y[2]};
    f.write(reinterpret_cast<const char *>(r), sizeof r);
  }
  return worst < 1e-9 ? 0 : 5;
}

// ---- ZMPQPWithConstraint (Wieber2006) through its class interface, and ql0001_ with the reference's signature ---------
static int test_wieber(const char *out)
{
  SimplePluginManager spm;
  CjrlHumanoidDynamicRobot robot(0.25, 0.14, 0.105);
  ZMPQPWithConstraint gen(&spm, "", &robot);
  { std::string m(":setpbwconstraint"); std::istringstream s("XY 0.04 0.04"); gen.CallMethod(m, s); }
  gen.SetSamplingPeriod(0.005); gen.SetTSingleSupport(0.78); gen.SetTDoubleSupport(0.02);
  { std::string m(":stepheight"); std::istringstream s("0.07"); gen.CallMethod(m, s); }
  std::deque<RelativeFootPosition> rel;
  const double seq[6][3] = {{0.0, -0.095, 0.0}, {0.2, 0.19, 0.0}, {0.2, -0.19, 0.0}, {0.2, 0.19, 0.0}, {0.2, -0.19, 0.0}, {0.0, 0.19, 0.0}};
  for (int i = 0; i < 6; ++i) {
    RelativeFootPosition r; std::memset(&r, 0, sizeof r);
    r.sx = seq[i][0]; r.sy = seq[i][1]; r.theta = seq[i][2]; r.SStime = 0.78; r.DStime = 0.02; r.stepType = 1;
    rel.push_back(r);
  }
  std::deque<ZMPPosition> zmp; std::deque<COMState> com; std::deque<FootAbsolutePosition> lf, rf;
  COMState start; S3Vector zstart;
  FootAbsolutePosition il, ir; std::memset(&il, 0, sizeof il); std::memset(&ir, 0, sizeof ir);
  il.x = 0.00949035; il.y = 0.095; ir.x = 0.00949035; ir.y = -0.095;
  gen.GetZMPDiscretization(zmp, com, rel, lf, rf, 0.0, start, zstart, il, ir);
  if (gen.LastStatus() != 0 || gen.PeriodsDone() < 100) { std::cerr << "status " << gen.LastStatus() << " periods " << gen.PeriodsDone() << std::endl; return 2; }
  // a small QP through ql0001_: min 1/2 |x|^2 - (1,1)x  s.t. x0 + x1 <= 1  ->  x = (0.5, 0.5), u = 0.5
  int m = 1, me = 0, mmax = 2, n = 2, nmax = 2, mnn = 5, iout = 0, ifail = -1, iprint = 0, lwar = 100, liwar = 10;
  double c[4] = {1, 0, 0, 1}, d[2] = {-1, -1}, a[4] = {-1, 0, -1, 0}, b[2] = {1, 0}, xl[2] = {-1e8, -1e8}, xu[2] = {1e8, 1e8};
  double x[2] = {0, 0}, u[5] = {0}, war[100], eps = 1e-8;
  int iwar[10] = {1};
  ql0001_(&m, &me, &mmax, &n, &nmax, &mnn, c, d, a, b, xl, xu, x, u, &iout, &ifail, &iprint, war, &lwar, iwar, &liwar, &eps);
  if (ifail != 0 || fabs(x[0] - 0.5) > 1e-12 || fabs(x[1] - 0.5) > 1e-12 || fabs(u[0] - 0.5) > 1e-12) {
    std::cerr << "ql0001_: ifail " << ifail << " x " << x[0] << " " << x[1] << " u " << u[0] << std::endl; return 3;
  }
  std::ofstream f(out, std::ios::binary);
  const double hdr[2] = {(double)com.size(), (double)gen.PeriodsDone()};
  f.write(reinterpret_cast<const char *>(hdr), sizeof hdr);
  for (size_t i = 0; i < com.size(); ++i) {
    const double r[8] = {com[i].x[0], com[i].x[1], com[i].x[2], com[i].y[0], com[i].y[1], com[i].y[2], zmp[i].px, zmp[i].py};
    f.write(reinterpret_cast<const char *>(r), sizeof r);
  }
  std::cout << "wieber: " << com.size() << " samples, " << gen.PeriodsDone() << " QP periods; ql0001_ ok" << std::endl;
  return 0;
}

int main(int argc, char **argv)
{
  try {
    std::string what = argc > 1 ? argv[1] : "";
    if (what == "optcholesky") return test_optcholesky();
    if (what == "herdt2010" && argc > 3) return test_herdt2010(argv[2], atoi(argv[3]), argc > 4 && std::string(argv[4]) == "emergency");
    if (what == "preview" && argc > 2) return test_preview(argv[2]);
    if (what == "kajita2003" && argc > 3) return test_kajita2003(argv[2], argv[3]);
    if (what == "preview1d" && argc > 4) return test_preview1d(argv[2], argv[3], argv[4]);
    if (what == "wieber" && argc > 2) return test_wieber(argv[2]);
    if (what == "twostage" && argc > 4) return test_twostage(argv[2], argv[3], atoi(argv[4]));
    if (what == "pldp" && argc > 3) return test_pldp(argv[2], argv[3]);
    if (what == "dimitrov" && argc > 2) return test_dimitrov(argv[2], argc > 3 && std::string(argv[3]) == "robust");
    std::cerr << "usage: host_api_test optcholesky | herdt2010 out.dat nticks | preview out.bin | pldp in.bin out.bin" << std::endl;
    return 64;
  } catch (const std::exception &e) {
    std::cerr << "exception: " << e.what() << std::endl;
    return 65;
  }
}
