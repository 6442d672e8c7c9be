/* oracle/oracle_preview.cpp - TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain restatement of the reference's Kajita2003 cart-table preview controller:
 *   gains : PreviewControl::ComputeOptimalWeights   src/PreviewControl/PreviewControl.cpp:198-322
 *           OptimalControllerSolver::ComputeWeights  src/PreviewControl/OptimalControllerSolver.cpp:200-352
 *   step  : PreviewControl::OneIterationOfPreview    src/PreviewControl/PreviewControl.cpp:324-374
 *           PreviewControl::OneIterationOfPreview1D  src/PreviewControl/PreviewControl.cpp:376-484
 *
 * Third-party arithmetic: the reference obtains the Riccati solution P from LAPACK dgges_ (ordered
 * generalized Schur form, OptimalControllerSolver.cpp:166-180) through jrl-mal, neither of which is
 * vendored.  P is the unique stabilising solution of the discrete algebraic Riccati equation
 *     P = A'PA - A'Pb (R + b'Pb)^-1 b'PA + c'Qc,
 * so it is restated here by the plain Riccati difference iteration run to a fixed point (slow but
 * obviously correct); tests/ pin it against scipy.linalg.solve_discrete_are and against the 5-digit
 * gains shipped in src/data/PreviewControlParameters.ini.
 *
 * Arithmetic order follows the reference statement by statement (uBLAS prod() = row-by-row sums
 * starting from 0, no FMA contraction on the reference's x86-64 -O3 build).
 */
#include <cmath>
#include <cstring>
#include <array>
#include <deque>
#include <thread>
#include <vector>

namespace {

const int MODE_WITHOUT_INITIALPOS = 1; /* OptimalControllerSolver.hh: MODE_WITHOUT_INITIALPOS */
const int MODE_WITH_INITIALPOS = 0;

/* n x n dense helpers (n <= 4), row-major. */
void matmul(const double *A, const double *B, double *C, int n, int m, int k)
{ /* C(n x k) = A(n x m) B(m x k) */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < k; ++j) {
      double s = 0.0;
      for (int l = 0; l < m; ++l) s += A[i * m + l] * B[l * k + j];
      C[i * k + j] = s;
    }
}

/* Riccati difference iteration to a fixed point.  A n x n, b n x 1, c 1 x n. */
int dare_fixed_point(const double *A, const double *b, const double *c, double Q, double R, int n,
                     double *P)
{
  std::vector<double> Pn(n * n), PA(n * n), AtPA(n * n), Pb(n), AtPb(n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) P[i * n + j] = c[i] * Q * c[j];
  for (int it = 0; it < 20000000; ++it) {
    matmul(P, A, PA.data(), n, n, n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int l = 0; l < n; ++l) s += A[l * n + i] * PA[l * n + j];
        AtPA[i * n + j] = s;
      }
    matmul(P, b, Pb.data(), n, n, 1);
    double bPb = 0.0;
    for (int i = 0; i < n; ++i) bPb += b[i] * Pb[i];
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int l = 0; l < n; ++l) s += A[l * n + i] * Pb[l];
      AtPb[i] = s;
    }
    double inv = 1.0 / (R + bPb);
    double diff = 0.0, norm = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double v = AtPA[i * n + j] - AtPb[i] * inv * AtPb[j] + c[i] * Q * c[j];
        diff = std::fmax(diff, std::fabs(v - P[i * n + j]));
        norm = std::fmax(norm, std::fabs(v));
        Pn[i * n + j] = v;
      }
    /* keep P symmetric */
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) P[i * n + j] = 0.5 * (Pn[i * n + j] + Pn[j * n + i]);
    if (diff <= 1e-15 * norm) return it;
  }
  return -1;
}

} // namespace

extern "C" {

/* PreviewControl::ComputeOptimalWeights (PreviewControl.cpp:198-322).
 * Outputs Ks, Kx[3], F[NL] with NL = (int)(preview_time / T) (PreviewControl.cpp:226), and the
 * cart-table matrices A (3x3 row-major), B[3], C[3] (PreviewControl.cpp:204-214).
 * Returns NL, or a negative value on failure. */
int oracle_preview_gains(double T, double preview_time, double zc, int mode, double *Ks, double *Kx,
                         double *F, int F_capacity, double *A_out, double *B_out, double *C_out)
{
  if (T == 0.0 || preview_time == 0.0) return -1; /* PreviewControl.cpp:220-224 */
  double A[9] = {1.0, T, T * T / 2.0, 0.0, 1.0, T, 0.0, 0.0, 1.0};
  double B[3] = {T * T * T / 6.0, T * T / 2.0, T};
  double C[3] = {1.0, 0.0, -zc / 9.81};
  if (A_out) std::memcpy(A_out, A, sizeof A);
  if (B_out) std::memcpy(B_out, B, sizeof B);
  if (C_out) std::memcpy(C_out, C, sizeof C);
  int Nl = (int)(preview_time / T);
  if (Nl > F_capacity) return -2;

  int n;
  double Ax[16], bx[4], cx[4], Q, R;
  if (mode == MODE_WITHOUT_INITIALPOS) { /* PreviewControl.cpp:229-282: augmented (e, dx) system */
    Q = 1.0;
    R = 1e-6;
    n = 4;
    std::memset(Ax, 0, sizeof Ax);
    double CA[3], CB = 0.0;
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int l = 0; l < 3; ++l) s += C[l] * A[l * 3 + j];
      CA[j] = s;
    }
    for (int l = 0; l < 3; ++l) CB += C[l] * B[l];
    Ax[0] = 1.0;
    for (int i = 0; i < 3; ++i) {
      Ax[0 * 4 + i + 1] = CA[i];
      for (int j = 0; j < 3; ++j) Ax[(i + 1) * 4 + j + 1] = A[i * 3 + j];
    }
    bx[0] = CB;
    for (int i = 0; i < 3; ++i) bx[i + 1] = B[i];
    cx[0] = 1.0;
    cx[1] = cx[2] = cx[3] = 0.0;
  } else { /* PreviewControl.cpp:284-303 */
    Q = 1.0;
    R = 1e-5;
    n = 3;
    std::memcpy(Ax, A, sizeof A);
    std::memcpy(bx, B, sizeof B);
    std::memcpy(cx, C, sizeof C);
  }

  double P[16];
  if (dare_fixed_point(Ax, bx, cx, Q, R, n, P) < 0) return -3;

  /* OptimalControllerSolver.cpp:303-349 */
  double Pb[4], bPb = 0.0;
  matmul(P, bx, Pb, n, n, 1);
  for (int i = 0; i < n; ++i) bPb += bx[i] * Pb[i];
  double la = 1.0 / (R + bPb);
  double PA[16], K[4];
  matmul(P, Ax, PA, n, n, n);
  for (int j = 0; j < n; ++j) {
    double s = 0.0;
    for (int l = 0; l < n; ++l) s += bx[l] * PA[l * n + j];
    K[j] = s * la;
  }
  /* BaseOfRecursion = (A - b K)^T ; Recursive0 = P c^T Q (WITHOUT_INITIALPOS) or c^T Q */
  double Base[16], Rec[4], Rec2[4];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Base[j * n + i] = Ax[i * n + j] - bx[i] * K[j];
  for (int i = 0; i < n; ++i) Rec[i] = cx[i] * Q;
  if (mode == MODE_WITHOUT_INITIALPOS) {
    matmul(P, Rec, Rec2, n, n, 1);
    std::memcpy(Rec, Rec2, sizeof(double) * n);
  }
  for (int k = 0; k < Nl; ++k) {
    double s = 0.0;
    for (int l = 0; l < n; ++l) s += la * bx[l] * Rec[l];
    F[k] = s;
    matmul(Base, Rec, Rec2, n, n, 1);
    std::memcpy(Rec, Rec2, sizeof(double) * n);
  }
  if (mode == MODE_WITHOUT_INITIALPOS) { /* PreviewControl.cpp:278-280 */
    *Ks = K[0];
    for (int i = 0; i < 3; ++i) Kx[i] = K[i + 1];
  } else { /* PreviewControl.cpp:297-301 */
    *Ks = K[0];
    for (int i = 0; i < 3; ++i) Kx[i] = K[i];
  }
  return Nl;
}

/* One axis of PreviewControl::OneIterationOfPreview (PreviewControl.cpp:336-369); `stride` is the
 * element stride of the ZMP reference (2 for interleaved px,py).  Statement order as the reference. */
static inline void preview_axis(const double *A, const double *B, const double *C, const double *Kx,
                                double Ks, const double *F, int NL, double *x, double *s,
                                const double *zmp, int stride, double *zmp_out, int simulation)
{
  double r = 0.0;
  for (int i = 0; i < 3; ++i) r += Kx[i] * x[i];
  double u = -r + Ks * (*s);
  for (int i = 0; i < NL; ++i) u += F[i] * zmp[(size_t)i * stride];
  double xn[3];
  for (int i = 0; i < 3; ++i) {
    double t = 0.0;
    for (int j = 0; j < 3; ++j) t += A[i * 3 + j] * x[j];
    xn[i] = t + u * B[i];
  }
  x[0] = xn[0];
  x[1] = xn[1];
  x[2] = xn[2];
  double z = 0.0;
  for (int i = 0; i < 3; ++i) z += C[i] * x[i];
  *zmp_out = z;
  if (simulation) *s += (zmp[0] - z);
}

/* PreviewControl::OneIterationOfPreview (PreviewControl.cpp:324-374): both axes, in place.
 * zmpref points at ZMPPositions[lindex] as interleaved (px,py) pairs. */
int oracle_preview_step(const double *A, const double *B, const double *C, const double *Kx, double Ks,
                        const double *F, int NL, double *x, double *y, double *sx, double *sy,
                        const double *zmpref_xy, int n_available, double *zmpx2, double *zmpy2,
                        int simulation)
{
  if (n_available < NL) return -1; /* LTHROW at PreviewControl.cpp:341-344 */
  preview_axis(A, B, C, Kx, Ks, F, NL, x, sx, zmpref_xy, 2, zmpx2, simulation);
  preview_axis(A, B, C, Kx, Ks, F, NL, y, sy, zmpref_xy + 1, 2, zmpy2, simulation);
  return 0;
}

/* Whole-trajectory driver: what ZMPPreviewControlWithMultiBodyZMP::FirstStageOfControl does with a
 * FIFO that is popped once per tick (ZMPPreviewControlWithMultiBodyZMP.cpp:378-446): step k uses the
 * window zmpref[k .. k+NL).  A trajectory of L samples yields L-NL+1 steps.
 * state = {x[3], y[3], sx, sy} in/out; com_out[k][6] = (x,dx,ddx,y,dy,ddy) after step k;
 * zmp_out[k][2].  Returns number of steps executed. */
int oracle_preview_run(const double *A, const double *B, const double *C, const double *Kx, double Ks,
                       const double *F, int NL, const double *zmpref_xy, int L, double *state,
                       double *com_out, double *zmp_out, int simulation)
{
  int nsteps = L - NL + 1;
  if (nsteps <= 0) return 0;
  double *x = state, *y = state + 3, *sx = state + 6, *sy = state + 7;
  for (int k = 0; k < nsteps; ++k) {
    double zx, zy;
    oracle_preview_step(A, B, C, Kx, Ks, F, NL, x, y, sx, sy, zmpref_xy + 2 * (size_t)k, L - k, &zx,
                        &zy, simulation);
    if (com_out) {
      double *c = com_out + 6 * (size_t)k;
      c[0] = x[0]; c[1] = x[1]; c[2] = x[2];
      c[3] = y[0]; c[4] = y[1]; c[5] = y[2];
    }
    if (zmp_out) {
      zmp_out[2 * (size_t)k] = zx;
      zmp_out[2 * (size_t)k + 1] = zy;
    }
  }
  return nsteps;
}

/* Batched ragged driver used by the CPU baseline: offsets[B+1] in samples (same packing as the
 * product's wg_preview_run_batch); outputs indexed like the inputs (row offsets[b]+k). */
long oracle_preview_run_batch(const double *A, const double *B, const double *C, const double *Kx,
                              double Ks, const double *F, int NL, int nb, const long long *offsets,
                              const double *zmpref_xy, double *state, double *com_out, double *zmp_out,
                              int simulation)
{
  long total = 0;
  for (int b = 0; b < nb; ++b) {
    long long o = offsets[b];
    int L = (int)(offsets[b + 1] - o);
    total += oracle_preview_run(A, B, C, Kx, Ks, F, NL, zmpref_xy + 2 * o, L, state + 8 * (size_t)b,
                                com_out ? com_out + 6 * o : nullptr, zmp_out ? zmp_out + 2 * o : nullptr,
                                simulation);
  }
  return total;
}

/* PreviewControl::OneIterationOfPreview1D, vector<double> overload with wrap-around
 * (PreviewControl.cpp:428-484). */
int oracle_preview_step_1d_wrap(const double *A, const double *B, const double *C, const double *Kx,
                                double Ks, const double *F, int NL, double *x, double *s,
                                const double *zmp, int size, int lindex, double *zmp_out, int simulation)
{
  if (size < NL) return -1; /* exit(0) in the reference, PreviewControl.cpp:449-454 */
  double r = 0.0;
  for (int i = 0; i < 3; ++i) r += Kx[i] * x[i];
  double u = -r + Ks * (*s);
  int TestSize = size - lindex - NL;
  if (TestSize >= 0) {
    for (int i = 0; i < NL; ++i) u += F[i] * zmp[lindex + i];
  } else {
    /* Reference quirk kept: first loop indexes F with the absolute index i (PreviewControl.cpp:466). */
    for (int i = lindex; i < size; ++i) u += F[i] * zmp[i];
    int still = NL - size + lindex;
    for (int i = 0; i < still; ++i) u += F[i] * zmp[i];
  }
  double xn[3];
  for (int i = 0; i < 3; ++i) {
    double t = 0.0;
    for (int j = 0; j < 3; ++j) t += A[i * 3 + j] * x[j];
    xn[i] = t + u * B[i];
  }
  x[0] = xn[0]; x[1] = xn[1]; x[2] = xn[2];
  double z = 0.0;
  for (int i = 0; i < 3; ++i) z += C[i] * x[i];
  *zmp_out = z;
  if (simulation) *s += (zmp[lindex] - z);
  return 0;
}


/* CPU-baseline driver (bench.py cpu_baseline / --impl reference): the reference's data layout and call
 * pattern - a std::deque of 48-byte ZMPPosition records (pgtypes.hh:90-104) read through operator[] at
 * lindex + i (PreviewControl.cpp:346-356) and popped once per tick by the caller
 * (ZMPPreviewControlWithMultiBodyZMP.cpp:437-438) - one walking problem per thread, `nthreads` host threads.
 * Results are identical to oracle_preview_run_batch (same statement order). */
struct RefZmpPosition { double px, py, pz, theta, time; int stepType; };

static void preview_walk_deque(const double *A, const double *B, const double *C, const double *Kx, double Ks,
                               const double *F, int NL, const double *zmpref_xy, int L, double *state,
                               double *com_out, double *zmp_out, int simulation)
{
  std::deque<RefZmpPosition> fifo;
  for (int k = 0; k < L; ++k) {
    RefZmpPosition z = {zmpref_xy[2 * (size_t)k], zmpref_xy[2 * (size_t)k + 1], 0.0, 0.0, 0.005 * k, 0};
    fifo.push_back(z);
  }
  double *x = state, *y = state + 3, *sx = state + 6, *sy = state + 7;
  int k = 0;
  while ((int)fifo.size() >= NL) {
    const unsigned lindex = 0;
    double rx = 0.0, ry = 0.0;
    for (int i = 0; i < 3; ++i) { rx += Kx[i] * x[i]; ry += Kx[i] * y[i]; }
    double ux = -rx + Ks * (*sx), uy = -ry + Ks * (*sy);
    for (int i = 0; i < NL; ++i) ux += F[i] * fifo[lindex + i].px;
    for (int i = 0; i < NL; ++i) uy += F[i] * fifo[lindex + i].py;
    double xn[3], yn[3];
    for (int i = 0; i < 3; ++i) {
      double t = 0.0, v = 0.0;
      for (int j = 0; j < 3; ++j) { t += A[i * 3 + j] * x[j]; v += A[i * 3 + j] * y[j]; }
      xn[i] = t + ux * B[i]; yn[i] = v + uy * B[i];
    }
    for (int i = 0; i < 3; ++i) { x[i] = xn[i]; y[i] = yn[i]; }
    double zx = 0.0, zy = 0.0;
    for (int i = 0; i < 3; ++i) { zx += C[i] * x[i]; zy += C[i] * y[i]; }
    if (simulation) { *sx += fifo[lindex].px - zx; *sy += fifo[lindex].py - zy; }
    if (com_out) { double *c = com_out + 6 * (size_t)k; c[0] = x[0]; c[1] = x[1]; c[2] = x[2]; c[3] = y[0]; c[4] = y[1]; c[5] = y[2]; }
    if (zmp_out) { zmp_out[2 * (size_t)k] = zx; zmp_out[2 * (size_t)k + 1] = zy; }
    fifo.pop_front();
    ++k;
  }
}

long oracle_preview_run_batch_mt(const double *A, const double *B, const double *C, const double *Kx,
                                 double Ks, const double *F, int NL, int nb, const long long *offsets,
                                 const double *zmpref_xy, double *state, double *com_out, double *zmp_out,
                                 int simulation, int nthreads)
{
  if (nthreads < 1) nthreads = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([=]() {
      for (int b = t; b < nb; b += nthreads) {
        long long o = offsets[b];
        int L = (int)(offsets[b + 1] - o);
        preview_walk_deque(A, B, C, Kx, Ks, F, NL, zmpref_xy + 2 * o, L, state + 8 * (size_t)b,
                           com_out ? com_out + 6 * o : nullptr, zmp_out ? zmp_out + 2 * o : nullptr, simulation);
      }
    });
  for (auto &t : th) t.join();
  long total = 0;
  for (int b = 0; b < nb; ++b) { long long L = offsets[b + 1] - offsets[b]; if (L >= NL) total += (long)(L - NL + 1); }
  return total;
}

/* ZMPPreviewControlWithMultiBodyZMP, the two-stage scheme with a caller-supplied multibody ZMP, restated with the
 * reference's FIFOs (src/PreviewControl/ZMPPreviewControlWithMultiBodyZMP.cpp):
 *   SetupFirstPhase (:530-600): FIFO = ZMPRefPositions[0 .. NL), PC1 = (start CoM, 0, 0), Delta = 0, all sums 0;
 *   SetupIterativePhase(i), i < NL (:602-664): FirstStageOfControl (preview at lindex 0 with Simulation = true,
 *       push the CoM, pop the FIFO front, :378-446), EvaluateMultiBodyZMP (delta = FIFO[0] - ZMPmultibody, :447-479),
 *       then push ZMPRefPositions[i + 1 + NL] - so ZMPRefPositions[NL] never enters the FIFO (:660);
 *   OneGlobalStepOfControl (:194-266): FirstStageOfControl, EvaluateMultiBodyZMP, SecondStageOfControl (preview on the
 *       delta FIFO at lindex 0, final CoM = FIFOCOM[0] + Delta on all three derivatives, pop both, :317-376); the caller
 *       then feeds the next reference sample (UpdateTheZMPRefQueue, :753).
 * zmb [n_zmb][2] is the multibody ZMP of first-stage tick k (the reference computes it from the posture realised for
 * that tick's CoM).  Outputs: stage1 [ticks][6], delta [ticks][2] (ticks = first-stage ticks run), final_com [steps][6];
 * returns steps = the number of OneGlobalStepOfControl calls the stream allows, or -1. */
long oracle_two_stage_run(const double *A, const double *B, const double *C, const double *Kx, double Ks,
                          const double *F, int NL, const double *zmpref_xy, long L, const double *zmb, long n_zmb,
                          const double *com_start_xy, double *stage1, double *delta, double *final_com, long *ticks_out)
{
  if (L < 2 * (long)NL + 1) return -1;
  std::deque<std::array<double, 2>> fifo_ref, fifo_delta;
  std::deque<std::array<double, 6>> fifo_com;
  for (int i = 0; i < NL; ++i) fifo_ref.push_back({zmpref_xy[2 * i], zmpref_xy[2 * i + 1]});
  double pc1x[3] = {com_start_xy[0], 0, 0}, pc1y[3] = {com_start_xy[1], 0, 0}, sx = 0, sy = 0;
  double dx[3] = {0, 0, 0}, dy[3] = {0, 0, 0}, sdx = 0, sdy = 0;
  std::vector<double> win(2 * (size_t)NL);
  long tick = 0;
  auto first_stage = [&]() {
    for (int i = 0; i < NL; ++i) { win[2 * i] = fifo_ref[i][0]; win[2 * i + 1] = fifo_ref[i][1]; }
    double zx, zy;
    oracle_preview_step(A, B, C, Kx, Ks, F, NL, pc1x, pc1y, &sx, &sy, win.data(), (int)fifo_ref.size(), &zx, &zy, 1);
    fifo_com.push_back({pc1x[0], pc1x[1], pc1x[2], pc1y[0], pc1y[1], pc1y[2]});
    if (stage1) for (int c = 0; c < 6; ++c) stage1[6 * tick + c] = fifo_com.back()[c];
    fifo_ref.pop_front();
  };
  auto evaluate_multibody = [&]() {
    std::array<double, 2> d = {fifo_ref[0][0] - zmb[2 * tick], fifo_ref[0][1] - zmb[2 * tick + 1]};
    fifo_delta.push_back(d);
    if (delta) { delta[2 * tick] = d[0]; delta[2 * tick + 1] = d[1]; }
  };
  for (int i = 0; i < NL; ++i) {               /* Setup */
    if (tick >= n_zmb) return -1;
    first_stage();
    evaluate_multibody();
    fifo_ref.push_back({zmpref_xy[2 * (size_t)(i + 1 + NL)], zmpref_xy[2 * (size_t)(i + 1 + NL) + 1]});
    ++tick;
  }
  long next_ref = 2 * (long)NL + 1, steps = 0;
  while ((long)fifo_ref.size() >= NL && tick < n_zmb) {   /* OneGlobalStepOfControl */
    first_stage();
    if (fifo_ref.empty()) break;
    evaluate_multibody();
    for (int i = 0; i < NL; ++i) { win[2 * i] = fifo_delta[i][0]; win[2 * i + 1] = fifo_delta[i][1]; }
    double zx, zy;
    oracle_preview_step(A, B, C, Kx, Ks, F, NL, dx, dy, &sdx, &sdy, win.data(), (int)fifo_delta.size(), &zx, &zy, 1);
    const std::array<double, 6> &c0 = fifo_com[0];
    if (final_com) {
      double *o = final_com + 6 * steps;
      o[0] = c0[0] + dx[0]; o[1] = c0[1] + dx[1]; o[2] = c0[2] + dx[2];
      o[3] = c0[3] + dy[0]; o[4] = c0[4] + dy[1]; o[5] = c0[5] + dy[2];
    }
    fifo_delta.pop_front();
    fifo_com.pop_front();
    ++steps; ++tick;
    if (next_ref < L) { fifo_ref.push_back({zmpref_xy[2 * next_ref], zmpref_xy[2 * next_ref + 1]}); ++next_ref; }
  }
  if (ticks_out) *ticks_out = tick;
  return steps;
}

} /* extern "C" */
