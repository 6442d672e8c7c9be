/* oracle/oracle_zmpdisc.cpp - TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain restatement of the reference's Kajita2003 front end, footsteps -> 5 ms ZMP reference + feet:
 *   step stack : StepStackHandler::PrepareForSupportFoot / CreateArcInStepStack /
 *                FinishOnTheLastCorrectSupportFoot          src/StepStackHandler.cpp:754-764, :299-457, :872-883
 *   reference  : ZMPDiscretization::InitOnLine              src/ZMPRefTrajectoryGeneration/ZMPDiscretization.cpp:319-513
 *                ZMPDiscretization::OnLineAddFoot           :573-1020
 *                ZMPDiscretization::EndPhaseOfTheWalking    :1129-1300
 *                ZMPDiscretization::FilterOutValues         :1046-1107   (window: InitializeFilter :233-257)
 *                ZMPDiscretization::UpdateCurrentSupportFootPosition :515-561
 *   feet       : FootTrajectoryGenerationStandard::UpdateFootPosition
 *                                                           src/FootTrajectoryGeneration/FootTrajectoryGenerationStandard.cpp:409-563
 *                Polynome3/4/5::SetParameters, Polynome::Compute
 *                                                           src/Mathematics/PolynomeFoot.cpp:38-56, :103-126, :177-200; Polynome.cpp:44-53
 *
 * PINNED by tests/golden/kajita_*.npz = columns 11-13, 20-25, 32-36 of the reference's four
 * tests/TestKajita2003*TestFGPI.datref.cmake files (feet and world ZMP reference; tests/test_zmpdisc.py).
 *
 * Statement order follows the reference (power-sum polynomial evaluation, uBLAS prod = sums from 0).
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <vector>

#include "../include/walkgen_b200.h"

namespace {

struct Poly {                 /* Polynome: r = sum c_i t^i evaluated with a running power (Polynome.cpp:44-53) */
  double c[6];
  int n;
  double eval(double t) const
  {
    double r = 0.0, pt = 1.0;
    for (int i = 0; i < n; ++i) {
      r += c[i] * pt;
      pt *= t;
    }
    return r;
  }
};

Poly poly3(double FT, double FP)
{ /* PolynomeFoot.cpp:38-56 */
  Poly p{};
  p.n = 4;
  double tmp = FT * FT;
  if (!(FP == 0.0 || FT == 0.0)) {
    p.c[2] = 3.0 * FP / tmp;
    p.c[3] = -2.0 * FP / (tmp * FT);
  }
  return p;
}

Poly poly4(double FT, double MP)
{ /* PolynomeFoot.cpp:103-126 */
  Poly p{};
  p.n = 5;
  double tmp = FT * FT;
  if (!(MP == 0.0 || tmp == 0.0)) {
    p.c[2] = 16.0 * MP / tmp;
    tmp = tmp * FT;
    p.c[3] = -32.0 * MP / tmp;
    tmp = tmp * FT;
    p.c[4] = 16.0 * MP / tmp;
  }
  return p;
}

Poly poly5(double FT, double FP)
{ /* PolynomeFoot.cpp:177-200 */
  Poly p{};
  p.n = 6;
  double tmp = FT * FT * FT;
  if (!(FP == 0.0 || tmp == 0.0)) {
    p.c[3] = 10 * FP / tmp;
    tmp *= FT;
    p.c[4] = -15 * FP / tmp;
    tmp *= FT;
    p.c[5] = 6 * FP / tmp;
  }
  return p;
}

struct Zmp {
  double px, py, theta;
  int type;
};
struct Foot {
  double x, y, z, theta, omega, omega2;
  int type;
};

struct Disc {
  wg_zmpdisc_params P;
  std::vector<double> window;
  double S[3][3], Sprev[3][3];          /* m_CurrentSupportFootPosition / m_PrevCurrentSupportFootPosition */
  double vpre[2];                        /* m_vdiffsupppre */
  double dTheta, dZmpTheta;              /* m_AngleDiffToSupportFootTheta / m_AngleDiffFromZMPThetaToSupportFootTheta */
  std::deque<wg_rel_step> rel;           /* m_RelativeFootPositions */
  std::vector<Zmp> fz;                   /* FinalZMPPositions */
  std::vector<Foot> fl, fr;              /* Final{Left,Right}FootAbsolutePositions */

  explicit Disc(const wg_zmpdisc_params &p) : P(p)
  { /* InitializeFilter, :233-257 */
    int n = (int)std::floor(P.filter_time / P.sampling_period);
    window.resize(n + 1);
    double sum = 0;
    for (int i = 0; i < n + 1; ++i) {
      double t = std::sin((M_PI * i) / n);
      window[i] = t * t;
    }
    for (int i = 0; i < n + 1; ++i) sum += window[i];
    for (int i = 0; i < n + 1; ++i) window[i] /= sum;
  }

  void filter_out(const std::vector<Zmp> &z, bool init)
  { /* FilterOutValues, :1046-1107 (pz is not carried: nothing downstream of this path reads it) */
    const int lshift = 2;
    for (int i = 0; i < (int)z.size(); ++i) {
      double a0 = 0, a1 = 0;
      int o = (int)fz.size() - 1 - lshift;
      for (int j = 0; j < (int)window.size(); ++j) {
        int r = i - j + lshift;
        if (r < 0) {
          if (init) {
            a0 += window[j] * z[lshift].px;
            a1 += window[j] * z[lshift].py;
          } else if (-r < o) {
            a0 += window[j] * fz[o + r].px;
            a1 += window[j] * fz[o + r].py;
          } else {
            a0 += window[j] * fz[0].px;
            a1 += window[j] * fz[0].py;
          }
        } else {
          if (r >= (int)z.size()) r = (int)z.size() - 1;
          a0 += window[j] * z[r].px;
          a1 += window[j] * z[r].py;
        }
      }
      fz.push_back(Zmp{a0, a1, z[i].theta, z[i].type});
    }
  }

  void update_support(const wg_rel_step &s)
  { /* UpdateCurrentSupportFootPosition, :515-561 */
    std::memcpy(Sprev, S, sizeof(S));
    double c = std::cos(s.theta * M_PI / 180.0), sn = std::sin(s.theta * M_PI / 180.0);
    double MM[2][2] = {{c, -sn}, {sn, c}}, O[2][2], N[2][2];
    for (int k = 0; k < 2; ++k)
      for (int l = 0; l < 2; ++l) O[k][l] = S[k][l];
    for (int k = 0; k < 2; ++k)
      for (int l = 0; l < 2; ++l) {
        double a = 0.0;
        for (int q = 0; q < 2; ++q) a += MM[k][q] * O[q][l];
        N[k][l] = a;
      }
    double v[2] = {s.sx, s.sy}, v2[2];
    for (int k = 0; k < 2; ++k) {
      double a = 0.0;
      for (int q = 0; q < 2; ++q) a += N[k][q] * v[q];
      v2[k] = a;
    }
    for (int k = 0; k < 2; ++k)
      for (int l = 0; l < 2; ++l) S[k][l] = N[k][l];
    for (int k = 0; k < 2; ++k) S[k][2] += v2[k];
  }

  void support_world(double w[2]) const
  { /* ZMPInWorldCoordinates = m_CurrentSupportFootPosition * (neutral, 1), :668-676 */
    double f[3] = {P.zmp_neutral[0], P.zmp_neutral[1], 1.0};
    for (int k = 0; k < 2; ++k) {
      double a = 0.0;
      for (int q = 0; q < 3; ++q) a += S[k][q] * f[q];
      w[k] = a;
    }
  }

  void who_is_support(const wg_rel_step &s, const Foot &L, const Foot &R, double zmp_theta, int *who)
  { /* :362-376 and :613-630 */
    if (s.sy < 0) {
      if (who) *who = -1;
      vpre[0] = R.x - L.x;
      vpre[1] = R.y - L.y;
      dTheta = R.theta - L.theta;
      dZmpTheta = R.theta - zmp_theta;
    } else {
      if (who) *who = 1;
      vpre[0] = -R.x + L.x;
      vpre[1] = -R.y + L.y;
      dTheta = L.theta - R.theta;
      dZmpTheta = L.theta - zmp_theta;
    }
  }

  void init_online(const wg_rel_step &first, const double *init_feet)
  { /* InitOnLine, :319-513 */
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) S[i][j] = (i == j) ? 1.0 : 0.0;
    std::memcpy(Sprev, S, sizeof(S));
    Foot L{init_feet[0], init_feet[1], 0.0, init_feet[2], 0.0, 0.0, 0};
    Foot R{init_feet[3], init_feet[4], 0.0, init_feet[5], 0.0, 0.0, 0};
    double zmp_theta = (R.theta + L.theta) / 2.0;
    who_is_support(first, L, R, zmp_theta, nullptr);
    int n = (int)(2 * P.preview_time / P.sampling_period);
    std::vector<Zmp> z(n);
    const double start[2] = {0.0, 0.0}, fin[2] = {P.zmp_neutral[0], P.zmp_neutral[1]};
    for (int i = 0; i < n; ++i) {
      double coef = (double)i / (double)n;
      z[i].px = start[0] + (fin[0] - start[0]) * coef;
      z[i].py = start[1] + (fin[1] - start[1]) * coef;
      z[i].theta = 0.0; /* CurrentAbsTheta, a local that stays 0 (:332) */
      z[i].type = 0;
      L.type = R.type = 10;
      fl.push_back(L);
      fr.push_back(R);
    }
    rel.clear();
    rel.push_back(first);
    filter_out(z, true);
  }

  int add_foot(const wg_rel_step &nw)
  { /* OnLineAddFoot, :573-1020 */
    const double T = P.sampling_period;
    Foot curL = fl.back(), curR = fr.back();
    double cur_zmp_theta = fz.back().theta;
    rel.push_back(nw);
    double lTdble = P.t_double, lTsingle = P.t_single;
    if (rel[1].ds_time != 0.0) {
      lTdble = rel[1].ds_time;
      lTsingle = rel[1].ss_time;
    }
    int who = 1;
    who_is_support(rel[0], curL, curR, cur_zmp_theta, &who);
    int add = (int)(unsigned)std::round((lTdble + lTsingle) / T);
    std::vector<Zmp> z(add, Zmp{0, 0, 0, 0});
    std::vector<Foot> l(add, Foot{0, 0, 0, 0, 0, 0, 0}), r(add, Foot{0, 0, 0, 0, 0, 0, 0});
    update_support(rel[0]);
    int n1 = (int)(unsigned)std::round(lTdble / T);
    double px0 = fz.back().px, py0 = fz.back().py, theta0 = fz.back().theta;
    double w[2];
    support_world(w);
    double dx = (w[0] - px0) / n1, dy = (w[1] - py0) / n1;
    const int t1 = rel[1].step_type;
    if (t1 == 3) {
      dx = (S[0][2] + P.zmp_shift[0] - px0) / n1;
      dy = (S[1][2] - py0) / n1;
    }
    if (t1 == 4) {
      dx = (S[0][2] + P.zmp_shift[2] - px0) / n1;
      dy = (S[1][2] - py0) / n1;
    }
    if (t1 == 5) {
      dx = (S[0][2] - (P.zmp_shift[0] + P.zmp_shift[2] + P.zmp_shift[1] + P.zmp_shift[3]) - px0) / n1;
      dy = (S[1][2] - py0) / n1;
    }
    int n2 = (int)(unsigned)std::round(lTsingle / T);
    if (n1 < 1 || n1 + n2 > add) return -1;
    int cur = 0;
    for (int k = 0; k < n1; ++k, ++cur) { /* double support */
      z[cur].px = px0 + k * dx;
      z[cur].py = py0 + k * dy;
      z[cur].theta = theta0;
      z[cur].type = t1 + 10;
      l[cur] = fl.back();
      l[cur].z = 0.0;
      r[cur] = fr.back();
      r[cur].z = 0.0;
      l[cur].type = r[cur].type = t1 + 10;
    }
    /* single support: where the swing foot goes */
    double next_theta = rel[1].theta;
    double rel_theta = next_theta + dTheta, rel_zmp_theta = next_theta + dZmpTheta;
    double c = std::cos(next_theta * M_PI / 180.0), s = std::sin(next_theta * M_PI / 180.0);
    double O[2][2] = {{c, -s}, {s, c}}, N[2][2];
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        double acc = 0.0;
        for (int q = 0; q < 2; ++q) acc += O[a][q] * S[q][b];
        N[a][b] = acc;
      }
    double v[2] = {rel[1].sx, rel[1].sy}, vd[2], vrel[2];
    for (int a = 0; a < 2; ++a) {
      double acc = 0.0;
      for (int q = 0; q < 2; ++q) acc += N[a][q] * v[q];
      vd[a] = acc;
    }
    vrel[0] = vd[0] + vpre[0];
    vrel[1] = vd[1] + vpre[1];
    vpre[0] = vd[0];
    vpre[1] = vd[1];
    double mod = lTsingle * P.modulation;
    double end_lift_param = (lTsingle - mod) * 0.5;
    Poly PX = poly5(mod, vrel[0]), PY = poly5(mod, vrel[1]), PZ = poly4(P.t_single, P.step_height);
    Poly PT = poly3(mod, rel_theta), PO = poly3(end_lift_param, P.omega), PO2 = poly3(mod, 2 * P.omega);
    Poly PZT = poly3(lTsingle, rel_zmp_theta);
    int init = cur - 1;
    double px02 = z[cur - 1].px, py02 = z[cur - 1].py;
    for (int k = 0; k < n2; ++k, ++cur) {
      support_world(w);
      z[cur].px = w[0];
      z[cur].py = w[1];
      if (t1 == 3 || t1 == 4) {
        double sh = (t1 == 3) ? P.zmp_shift[1] : P.zmp_shift[3];
        dx = (S[0][2] + sh - px02) / n2;
        dy = (S[1][2] - py02) / n2;
        z[cur].px = z[cur - 1].px + dx;
        z[cur].py = z[cur - 1].py + dy;
      }
      z[cur].theta = PZT.eval(k * T) + z[init].theta;
      z[cur].type = who * rel[0].step_type;
      /* UpdateFootPosition, FootTrajectoryGenerationStandard.cpp:409-563 */
      std::vector<Foot> &sup = (who == 1) ? l : r, &swg = (who == 1) ? r : l;
      unsigned kk = cur - init;
      double lt = kk * T;
      double end_lift = (P.t_single - mod) * 0.5, start_land = end_lift + mod;
      sup[cur] = sup[cur - 1];
      sup[cur].type = -t1;
      Foot &f = swg[cur];
      const Foot &f0 = swg[init];
      f.type = t1;
      if (lt < end_lift) {
        f.x = f0.x;
        f.y = f0.y;
        f.theta = f0.theta;
      } else if (lt < start_land) {
        f.x = f0.x + PX.eval(lt - end_lift);
        f.y = f0.y + PY.eval(lt - end_lift);
        f.theta = f0.theta + PT.eval(lt - end_lift);
      } else {
        f.x = f0.x + PX.eval(mod);
        f.y = f0.y + PY.eval(mod);
        f.theta = f0.theta + PT.eval(mod);
      }
      f.z = f0.z + PZ.eval(lt);
      if (lt < end_lift)
        f.omega = PO.eval(lt);
      else if (lt < start_land)
        f.omega = P.omega - PO2.eval(lt - end_lift);
      else
        f.omega = PO.eval(lt - start_land) - P.omega;
      double lo = f.omega * M_PI / 180.0, lth = f.theta * M_PI / 180.0;
      double cc = std::cos(lth), ss = std::sin(lth);
      double dX, dFZ;
      const double B = P.foot_b, H = P.foot_h, F = P.foot_f;
      if (lo < 0) {
        double X1 = B * std::cos(-lo), X2 = H * std::sin(-lo), Z1 = H * std::cos(-lo), Z2 = B * std::sin(-lo);
        dX = -(B - X1 + X2);
        dFZ = Z1 + Z2 - H;
      } else {
        double X1 = F * std::cos(lo), X2 = H * std::sin(lo), Z1 = H * std::cos(lo), Z2 = F * std::sin(lo);
        dX = (F - X1 + X2);
        dFZ = Z1 + Z2 - H;
      }
      f.x += cc * dX;
      f.y += ss * dX;
      f.z += dFZ;
    }
    rel.pop_front();
    for (int i = 0; i < add; ++i) {
      fl.push_back(l[i]);
      fr.push_back(r[i]);
    }
    filter_out(z, false);
    return 0;
  }

  int end_phase()
  { /* EndPhaseOfTheWalking, :1129-1300 */
    const double T = P.sampling_period;
    if (!rel.empty()) update_support(rel[0]);
    int n = (int)(unsigned)std::round(P.t_double / (2 * T));
    if (n < 1) return -1;
    int tail = (int)(3.0 * P.preview_time / T);
    std::vector<Zmp> z(n + tail);
    double px0 = fz.back().px, py0 = fz.back().py;
    double pxf = 0.5 * (S[0][2] + Sprev[0][2]), pyf = 0.5 * (S[1][2] + Sprev[1][2]);
    double dx = (pxf - px0) / (double)n, dy = (pyf - py0) / (double)n;
    z[0] = Zmp{px0 + dx, py0 + dy, fz.back().theta, 0};
    for (int k = 1; k < n + tail; ++k) {
      if (k < n)
        z[k] = Zmp{z[k - 1].px + dx, z[k - 1].py + dy, z[k - 1].theta, 0};
      else
        z[k] = Zmp{z[k - 1].px, z[k - 1].py, z[k - 1].theta, 0};
    }
    for (int k = 0; k < n + tail; ++k) {
      Foot L = fl.back(), R = fr.back();
      L.type = R.type = 0;
      fl.push_back(L);
      fr.push_back(R);
    }
    filter_out(z, false);
    return 0;
  }
};

int push_step(wg_rel_step *steps, int cap, int *n, const wg_rel_step &s)
{
  if (*n >= cap) return -1;
  steps[(*n)++] = s;
  return 0;
}

} /* namespace */

extern "C" {

void oracle_zmpdisc_default_params(wg_zmpdisc_params *p)
{ /* tests/CommonTools.cpp:56-69 (only the first nine commands are sent: loop bound 9, :71) */
  std::memset(p, 0, sizeof(*p));
  p->sampling_period = 0.005;
  p->preview_time = 1.6;
  p->t_single = 0.78;
  p->t_double = 0.02;
  p->step_height = 0.07;
  p->omega = 0.0;
  p->modulation = 0.9;
  p->filter_time = 0.05;
  p->foot_h = 0.105; /* only used when omega != 0 */
  p->foot_b = 0.1;
  p->foot_f = 0.13;
}

/* StepStackHandler::PrepareForSupportFoot, StepStackHandler.cpp:754-764 */
int oracle_steps_support_foot(wg_rel_step *steps, int cap, int *n, int support_foot, double ss, double ds)
{
  wg_rel_step s{};
  s.sx = 0;
  s.sy = support_foot * 0.095;
  s.theta = 0;
  s.ss_time = ss;
  s.ds_time = ds;
  s.step_type = 1; /* uninitialised in the reference; any value other than 3/4/5 behaves the same */
  return push_step(steps, cap, n, s);
}

/* StepStackHandler::CreateArcInStepStack, StepStackHandler.cpp:299-457 */
int oracle_steps_arc(wg_rel_step *steps, int cap, int *n, double x, double y, double arc_deg, int support_foot,
                     double ss, double ds, int *keep_last)
{
  double StepMax = 0.15;
  double OmegaTotal = arc_deg * M_PI / 180.0;
  int Dir = -1;
  double R = std::sqrt(x * x + y * y);
  double nf = OmegaTotal * R / StepMax;
  int N = (int)std::floor(nf);
  double LastStep = OmegaTotal * R - N * StepMax;
  double OmegaStep = StepMax / R;
  double LastOmegaStep = OmegaTotal - OmegaStep * N;
  OmegaStep = OmegaStep * 180.0 / M_PI;
  LastOmegaStep = LastOmegaStep * 180.0 / M_PI;
  if (x < 0) {
    StepMax = -StepMax;
    LastStep = -LastStep;
    Dir = 1;
  }
  if (y < 0) {
    OmegaStep = -OmegaStep;
    LastOmegaStep = -LastOmegaStep;
  }
  double Omegak = 0.0, Omegakp = 0.0;
  int SF = support_foot;
  for (int i = 0; i < N + 1; ++i) {
    double dO = OmegaStep;
    if (i == N) {
      if (LastStep == 0.0) break;
      dO = LastOmegaStep;
    }
    Omegakp = Omegak;
    Omegak = Omegak + dO;
    double c = std::cos(Omegak * M_PI / 180.0), s = std::sin(Omegak * M_PI / 180.0);
    double cp = std::cos(Omegakp * M_PI / 180.0), sp = std::sin(Omegakp * M_PI / 180.0);
    double lv0 = (R + Dir * SF * 0.095) * s - (R - Dir * SF * 0.095) * sp;
    double lv1 = -((R + Dir * SF * 0.095) * c - (R - Dir * SF * 0.095) * cp);
    wg_rel_step st{};
    st.sx = (0.0 + c * lv0) + s * lv1; /* A = [[c, s], [-s, c]] */
    st.sy = (0.0 + -s * lv0) + c * lv1;
    st.theta = dO;
    st.ss_time = ss;
    st.ds_time = ds;
    st.step_type = 1;
    if (push_step(steps, cap, n, st)) return -1;
    SF = -SF;
  }
  if (keep_last) *keep_last = SF;
  return 0;
}

/* StepStackHandler::FinishOnTheLastCorrectSupportFoot, StepStackHandler.cpp:872-883 */
int oracle_steps_last_support(wg_rel_step *steps, int cap, int *n, int keep_last, double ss, double ds)
{
  wg_rel_step s{};
  s.sx = 0;
  s.sy = keep_last * 0.19;
  s.theta = 0;
  s.ss_time = ss;
  s.ds_time = ds;
  s.step_type = 0;
  return push_step(steps, cap, n, s);
}

/* ZMPDiscretization::GetZMPDiscretization (:145-175).  Returns the number of samples L, or < 0.
 * zmp [L][3] = (px, py, theta); left/right [L][6] = (x, y, z, theta, omega, omega2); types [L][3]. */
long oracle_zmpdisc_run(const wg_zmpdisc_params *p, int n_steps, const wg_rel_step *steps, const double *init_feet,
                        long cap, double *zmp, double *left, double *right, int32_t *types)
{
  if (n_steps < 1) return -1;
  Disc d(*p);
  d.init_online(steps[0], init_feet);
  for (int i = 1; i < n_steps; ++i)
    if (d.add_foot(steps[i])) return -2;
  if (d.end_phase()) return -3;
  long L = (long)d.fz.size();
  if (L > cap) return -4;
  for (long i = 0; i < L; ++i) {
    if (zmp) {
      zmp[3 * i] = d.fz[i].px;
      zmp[3 * i + 1] = d.fz[i].py;
      zmp[3 * i + 2] = d.fz[i].theta;
    }
    const Foot *f[2] = {&d.fl[i], &d.fr[i]};
    double *o[2] = {left, right};
    for (int s = 0; s < 2; ++s)
      if (o[s]) {
        double *q = o[s] + 6 * i;
        q[0] = f[s]->x; q[1] = f[s]->y; q[2] = f[s]->z; q[3] = f[s]->theta; q[4] = f[s]->omega; q[5] = f[s]->omega2;
      }
    if (types) {
      types[3 * i] = d.fz[i].type;
      types[3 * i + 1] = d.fl[i].type;
      types[3 * i + 2] = d.fr[i].type;
    }
  }
  return L;
}

} /* extern "C" */
