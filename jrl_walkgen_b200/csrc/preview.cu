// preview.cu - batched Kajita2003 cart-table preview control for sm_100a.
//
// Replaces PreviewControl::OneIterationOfPreview (src/PreviewControl/PreviewControl.cpp:324-374) and
// its 1-D variants (:376-484) for whole ragged batches of trajectories, and the host-side gain
// computation PreviewControl::ComputeOptimalWeights (:198-322) / OptimalControllerSolver::ComputeWeights
// (src/PreviewControl/OptimalControllerSolver.cpp:200-352).
//
// B200-first restructuring.  The reference evaluates, per 5 ms tick and per axis,
//     u = -Kx.x + Ks.s + sum_{i<NL} F[i] * p[k+i]          (640 MACs through a deque of 48-byte structs)
// and then the 3-state update.  The preview sum does not depend on the state, so it is a FIR filter
// of the ZMP reference and is separated from the recursion:
//   kernel 1  preview_fir_kernel   : fir[k] = sum_i F[i] p[k+i] for every step of every trajectory.
//             This is >94% of the flops (1280 of 1360 per step).  Each thread produces R=8 consecutive
//             outputs for both axes from a register-resident sliding window: per tap it issues ONE
//             128-bit shared-memory load and 16 DFMAs, so the FP64 pipe (64 DFMA/clk/SM), not the LSU,
//             is the limiter.  The ZMP tile is staged once per block in shared memory (coalesced
//             128-bit global loads) with a 9/8 padding so that the stride-8 per-thread windows are
//             bank-conflict free.  HBM traffic is the streaming minimum (each sample read once per
//             tile + 30% halo, fir written once).
//   kernel 2  preview_recur_kernel : the 4-state (x, dx, ddx, s) recursion, one thread per
//             (trajectory, axis), in the reference's statement order.
//
#include "wg_common.h"
#include <vector>
#include <cmath>
#include <algorithm>

// ---------------------------------------------------------------------------------------------
// Host: gains by structure-preserving doubling (SDA) for the DARE
//     P = A'PA - A'Pb (R + b'Pb)^-1 b'PA + c'Qc
// ---------------------------------------------------------------------------------------------
namespace {

struct Mat {  // tiny dense n x n (n <= 4), row-major
  int n;
  double a[16];
  double &operator()(int i, int j) { return a[i * n + j]; }
  double operator()(int i, int j) const { return a[i * n + j]; }
};

Mat mm(const Mat &A, const Mat &B)
{
  Mat C{A.n, {0}};
  for (int i = 0; i < A.n; ++i)
    for (int j = 0; j < A.n; ++j) {
      double s = 0;
      for (int k = 0; k < A.n; ++k) s += A(i, k) * B(k, j);
      C(i, j) = s;
    }
  return C;
}
Mat tr(const Mat &A)
{
  Mat C{A.n, {0}};
  for (int i = 0; i < A.n; ++i)
    for (int j = 0; j < A.n; ++j) C(i, j) = A(j, i);
  return C;
}
Mat add(const Mat &A, const Mat &B)
{
  Mat C{A.n, {0}};
  for (int i = 0; i < A.n * A.n; ++i) C.a[i] = A.a[i] + B.a[i];
  return C;
}
// X = W^-1 B by Gaussian elimination with partial pivoting.
bool solve(Mat W, Mat B, Mat &X)
{
  int n = W.n;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(W(r, c)) > std::fabs(W(piv, c))) piv = r;
    if (W(piv, c) == 0.0) return false;
    if (piv != c)
      for (int j = 0; j < n; ++j) { std::swap(W(piv, j), W(c, j)); std::swap(B(piv, j), B(c, j)); }
    for (int r = c + 1; r < n; ++r) {
      double f = W(r, c) / W(c, c);
      for (int j = c; j < n; ++j) W(r, j) -= f * W(c, j);
      for (int j = 0; j < n; ++j) B(r, j) -= f * B(c, j);
    }
  }
  X.n = n;
  for (int j = 0; j < n; ++j)
    for (int r = n - 1; r >= 0; --r) {
      double s = B(r, j);
      for (int k = r + 1; k < n; ++k) s -= W(r, k) * X(k, j);
      X(r, j) = s / W(r, r);
    }
  return true;
}

bool dare_sda(const Mat &A0, const double *b, const double *c, double Q, double R, Mat &P)
{
  int n = A0.n;
  Mat A = A0, G{n, {0}}, H{n, {0}}, I{n, {0}};
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      G(i, j) = b[i] * b[j] / R;
      H(i, j) = c[i] * Q * c[j];
      I(i, j) = (i == j);
    }
  for (int it = 0; it < 200; ++it) {
    Mat W = add(I, mm(G, H));
    Mat WiA, WiG;
    if (!solve(W, A, WiA)) return false;   // W^-1 A
    if (!solve(W, G, WiG)) return false;   // W^-1 G
    Mat At = tr(A);
    Mat A1 = mm(A, WiA);
    Mat G1 = add(G, mm(mm(A, WiG), At));
    Mat H1 = add(H, mm(mm(At, H), WiA));
    double diff = 0, norm = 0;
    for (int i = 0; i < n * n; ++i) {
      diff = std::fmax(diff, std::fabs(H1.a[i] - H.a[i]));
      norm = std::fmax(norm, std::fabs(H1.a[i]));
    }
    A = A1; G = G1; H = H1;
    if (diff <= 1e-16 * norm) break;
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) P(i, j) = 0.5 * (H(i, j) + H(j, i));
  P.n = n;
  return true;
}

}  // namespace

extern "C" int wg_preview_gains(double T, double preview_time, double zc, int mode, wg_preview_gains_t *out)
{
  if (!out || T <= 0.0 || preview_time <= 0.0) return WG_ERR_INVALID;
  int NL = (int)(preview_time / T);
  if (NL <= 0 || NL > WG_PREVIEW_MAX_NL) return WG_ERR_INVALID;
  std::memset(out, 0, sizeof *out);
  out->T = T; out->preview_time = preview_time; out->zc = zc; out->mode = mode; out->NL = NL;
  const double A[9] = {1.0, T, T * T / 2.0, 0.0, 1.0, T, 0.0, 0.0, 1.0};
  const double B[3] = {T * T * T / 6.0, T * T / 2.0, T};
  const double C[3] = {1.0, 0.0, -zc / 9.81};
  std::memcpy(out->A, A, sizeof A);
  std::memcpy(out->B, B, sizeof B);
  std::memcpy(out->C, C, sizeof C);

  Mat Ax{0, {0}};
  double bx[4] = {0}, cx[4] = {0}, Q = 1.0, R;
  if (mode == WG_PREVIEW_MODE_WITHOUT_INITIALPOS) {
    // augmented (integrated error, state increment) system, PreviewControl.cpp:237-262
    R = 1e-6;
    Ax.n = 4;
    Ax(0, 0) = 1.0;
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int l = 0; l < 3; ++l) s += C[l] * A[l * 3 + j];
      Ax(0, j + 1) = s;
      for (int i = 0; i < 3; ++i) Ax(i + 1, j + 1) = A[i * 3 + j];
    }
    for (int l = 0; l < 3; ++l) { bx[0] += C[l] * B[l]; bx[l + 1] = B[l]; }
    cx[0] = 1.0;
  } else if (mode == WG_PREVIEW_MODE_WITH_INITIALPOS) {
    R = 1e-5;
    Ax.n = 3;
    for (int i = 0; i < 9; ++i) Ax.a[i] = A[i];
    for (int i = 0; i < 3; ++i) { bx[i] = B[i]; cx[i] = C[i]; }
  } else {
    return WG_ERR_INVALID;
  }
  int n = Ax.n;
  Mat P{n, {0}};
  if (!dare_sda(Ax, bx, cx, Q, R, P)) return WG_ERR_INVALID;

  double Pb[4] = {0}, bPb = 0;
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) Pb[i] += P(i, j) * bx[j];
  }
  for (int i = 0; i < n; ++i) bPb += bx[i] * Pb[i];
  const double la = 1.0 / (R + bPb);
  Mat PA = mm(P, Ax);
  double K[4] = {0};
  for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int l = 0; l < n; ++l) s += bx[l] * PA(l, j);
    K[j] = s * la;
  }
  // F[k] = la b' ((A - bK)')^k (P c'Q | c'Q)
  double rec[4], nxt[4];
  for (int i = 0; i < n; ++i) rec[i] = cx[i] * Q;
  if (mode == WG_PREVIEW_MODE_WITHOUT_INITIALPOS) {
    for (int i = 0; i < n; ++i) { nxt[i] = 0; for (int j = 0; j < n; ++j) nxt[i] += P(i, j) * rec[j]; }
    std::memcpy(rec, nxt, sizeof rec);
  }
  for (int k = 0; k < NL; ++k) {
    double s = 0;
    for (int l = 0; l < n; ++l) s += la * bx[l] * rec[l];
    out->F[k] = s;
    for (int i = 0; i < n; ++i) {
      nxt[i] = 0;
      for (int j = 0; j < n; ++j) nxt[i] += (Ax(j, i) - bx[j] * K[i]) * rec[j];
    }
    std::memcpy(rec, nxt, sizeof rec);
  }
  out->Ks = K[0];
  if (mode == WG_PREVIEW_MODE_WITHOUT_INITIALPOS)
    for (int i = 0; i < 3; ++i) out->Kx[i] = K[i + 1];
  else
    for (int i = 0; i < 3; ++i) out->Kx[i] = K[i];
  return WG_OK;
}

// ---------------------------------------------------------------------------------------------
// Device side
// ---------------------------------------------------------------------------------------------
constexpr int FIR_R = 8;                      // outputs per thread
constexpr int FIR_THREADS = 128;
constexpr int FIR_TILE = FIR_R * FIR_THREADS; // outputs per block

struct PreviewConsts {
  double A[9], B[3], C[3], Kx[3], Ks;
  int NL, NLpad;
};
__constant__ PreviewConsts c_pc;
__constant__ double c_F[WG_PREVIEW_MAX_NL + 8];

struct FirTile { int traj; int start; };      // outputs [start, start+FIR_TILE) of trajectory traj

struct wg_preview_plan {
  wg_ctx *ctx;
  int B;
  int NL;
  int64_t total_samples, total_steps;
  int n_tiles;
  int64_t *d_offsets;
  FirTile *d_tiles;
  double2 *d_fir;      // [total_samples]
  // staging buffers for WG_MEM_HOST calls
  double *d_zmp, *d_state, *d_com, *d_zmpout;
};

__device__ __forceinline__ int pad9(int e) { return e + (e >> 3); }

// fir[o+k] = sum_{i<NL} F[i] * p[o+k+i]  for k in the tile, both axes.
__global__ void __launch_bounds__(FIR_THREADS)
preview_fir_kernel(const FirTile *__restrict__ tiles, const int64_t *__restrict__ offsets,
                   const double2 *__restrict__ p, double2 *__restrict__ fir)
{
  extern __shared__ double2 sp[];             // padded tile of (px,py)
  const FirTile tile = tiles[blockIdx.x];
  const int64_t o = offsets[tile.traj];
  const int L = (int)(offsets[tile.traj + 1] - o);
  const int NL = c_pc.NL, NLpad = c_pc.NLpad;
  const int nsteps = L - NL + 1;
  const int span = FIR_TILE + NLpad;           // samples needed by this tile
  const double2 *src = p + o + tile.start;
  const int avail = L - tile.start;            // samples that exist from tile.start on
  for (int e = threadIdx.x; e < span; e += FIR_THREADS) {
    double2 v = make_double2(0.0, 0.0);
    if (e < avail) v = __ldg(src + e);
    sp[pad9(e)] = v;
  }
  __syncthreads();

  const int t = threadIdx.x;
  double ax[FIR_R], ay[FIR_R], wx[FIR_R], wy[FIR_R];
#pragma unroll
  for (int r = 0; r < FIR_R; ++r) {
    ax[r] = 0.0; ay[r] = 0.0;
    double2 v = sp[pad9(FIR_R * t + r)];
    wx[r] = v.x; wy[r] = v.y;
  }
  // window invariant at tap j: w[(j+r) % 8] holds p[8t + r + j]
  const double2 *wp = sp + pad9(FIR_R * t + FIR_R);   // next sample to enter the window
  for (int jj = 0; jj < NLpad; jj += FIR_R) {
#pragma unroll
    for (int u = 0; u < FIR_R; ++u) {
      const double f = c_F[jj + u];
#pragma unroll
      for (int r = 0; r < FIR_R; ++r) {
        ax[r] = fma(f, wx[(u + r) % FIR_R], ax[r]);
        ay[r] = fma(f, wy[(u + r) % FIR_R], ay[r]);
      }
      double2 v = wp[u];                        // p[8t + 8 + jj + u]; 8-aligned group => no pad inside
      wx[u] = v.x; wy[u] = v.y;
    }
    wp += FIR_R + 1;                            // 8 samples + 1 padding slot
  }
  const int k0 = tile.start + FIR_R * t;
  double2 *dst = fir + o + k0;
#pragma unroll
  for (int r = 0; r < FIR_R; ++r)
    if (k0 + r < nsteps) dst[r] = make_double2(ax[r], ay[r]);
}

// One thread per (trajectory, axis): the recursion of OneIterationOfPreview in statement order.
__global__ void __launch_bounds__(64)
preview_recur_kernel(int B, const int64_t *__restrict__ offsets, const double2 *__restrict__ p,
                     const double2 *__restrict__ fir, double *__restrict__ state,
                     double *__restrict__ com, double *__restrict__ zmp, int simulation)
{
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = tid >> 1, axis = tid & 1;
  if (b >= B) return;
  const int64_t o = offsets[b];
  const int L = (int)(offsets[b + 1] - o);
  const int nsteps = L - c_pc.NL + 1;
  if (nsteps <= 0) return;
  const double A01 = c_pc.A[1], A02 = c_pc.A[2], A12 = c_pc.A[5];
  const double B0 = c_pc.B[0], B1 = c_pc.B[1], B2 = c_pc.B[2];
  const double C0 = c_pc.C[0], C1 = c_pc.C[1], C2 = c_pc.C[2];
  const double K0 = c_pc.Kx[0], K1 = c_pc.Kx[1], K2 = c_pc.Kx[2], Ks = c_pc.Ks;
  double *st = state + 8 * (size_t)b;
  double x0 = st[3 * axis + 0], x1 = st[3 * axis + 1], x2 = st[3 * axis + 2], s = st[6 + axis];
  const double *pf = reinterpret_cast<const double *>(fir + o) + axis;
  const double *pp = reinterpret_cast<const double *>(p + o) + axis;
  double *pc = com ? com + 6 * o + 3 * axis : nullptr;
  double *pz = zmp ? zmp + 2 * o + axis : nullptr;
#pragma unroll 4
  for (int k = 0; k < nsteps; ++k) {
    const double f = __ldg(pf + 2 * (size_t)k);
    const double pk = __ldg(pp + 2 * (size_t)k);
    double r = K0 * x0;
    r = fma(K1, x1, r);
    r = fma(K2, x2, r);
    double u = fma(Ks, s, -r) + f;
    // x = A x + u B   (A = [[1,T,T^2/2],[0,1,T],[0,0,1]])
    double n0 = fma(A02, x2, fma(A01, x1, x0));
    double n1 = fma(A12, x2, x1);
    x0 = fma(u, B0, n0);
    x1 = fma(u, B1, n1);
    x2 = fma(u, B2, x2);
    double z = fma(C2, x2, fma(C1, x1, C0 * x0));
    if (simulation) s += (pk - z);
    if (pc) { pc[6 * (size_t)k] = x0; pc[6 * (size_t)k + 1] = x1; pc[6 * (size_t)k + 2] = x2; }
    if (pz) pz[2 * (size_t)k] = z;
  }
  st[3 * axis + 0] = x0; st[3 * axis + 1] = x1; st[3 * axis + 2] = x2; st[6 + axis] = s;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int wg_preview_set_gains(wg_ctx *ctx, const wg_preview_gains_t *g)
{
  if (!ctx || !g || g->NL <= 0 || g->NL > WG_PREVIEW_MAX_NL) return WG_ERR_INVALID;
  wg_device_guard guard(ctx->device);
  PreviewConsts pc;
  std::memcpy(pc.A, g->A, sizeof pc.A);
  std::memcpy(pc.B, g->B, sizeof pc.B);
  std::memcpy(pc.C, g->C, sizeof pc.C);
  std::memcpy(pc.Kx, g->Kx, sizeof pc.Kx);
  pc.Ks = g->Ks;
  pc.NL = g->NL;
  pc.NLpad = (g->NL + FIR_R - 1) / FIR_R * FIR_R;
  std::vector<double> F(WG_PREVIEW_MAX_NL + 8, 0.0);
  std::copy(g->F, g->F + g->NL, F.begin());
  WG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  WG_CUDA(ctx, cudaMemcpyToSymbol(c_pc, &pc, sizeof pc));
  WG_CUDA(ctx, cudaMemcpyToSymbol(c_F, F.data(), sizeof(double) * F.size()));
  ctx->preview_gains = *g;
  ctx->preview_ready = true;
  return WG_OK;
}

int wg_preview_plan_create(wg_ctx *ctx, int B, const int64_t *offsets, wg_preview_plan **out)
{
  if (!ctx || !out || B < 0 || (B > 0 && !offsets)) return WG_ERR_INVALID;
  if (!ctx->preview_ready) return wg_fail(ctx, WG_ERR_NOT_READY, "wg_preview_set_gains not called");
  wg_device_guard guard(ctx->device);
  *out = nullptr;
  const int NL = ctx->preview_gains.NL;
  std::vector<FirTile> tiles;
  int64_t total_steps = 0;
  for (int b = 0; b < B; ++b) {
    int64_t L = offsets[b + 1] - offsets[b];
    if (L < 0 || L > 0x3fffffff) return wg_fail(ctx, WG_ERR_INVALID, "offsets must be non-decreasing");
    int nsteps = (int)L - NL + 1;
    if (nsteps <= 0) continue;
    total_steps += nsteps;
    for (int s = 0; s < nsteps; s += FIR_TILE) tiles.push_back(FirTile{b, s});
  }
  wg_preview_plan *pl = new (std::nothrow) wg_preview_plan();
  if (!pl) return WG_ERR_ALLOC;
  std::memset(pl, 0, sizeof *pl);
  pl->ctx = ctx; pl->B = B; pl->NL = NL;
  pl->total_samples = B > 0 ? offsets[B] - offsets[0] : 0;
  pl->total_steps = total_steps;
  pl->n_tiles = (int)tiles.size();
  if (B > 0 && offsets[0] != 0) { delete pl; return wg_fail(ctx, WG_ERR_INVALID, "offsets[0] must be 0"); }
  cudaError_t e = cudaMalloc(&pl->d_offsets, sizeof(int64_t) * (B + 1));
  if (e == cudaSuccess) e = cudaMalloc(&pl->d_tiles, sizeof(FirTile) * std::max<size_t>(1, tiles.size()));
  if (e == cudaSuccess) e = cudaMalloc(&pl->d_fir, sizeof(double2) * std::max<int64_t>(1, pl->total_samples));
  if (e == cudaSuccess && B > 0)
    e = cudaMemcpyAsync(pl->d_offsets, offsets, sizeof(int64_t) * (B + 1), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && !tiles.empty())
    e = cudaMemcpyAsync(pl->d_tiles, tiles.data(), sizeof(FirTile) * tiles.size(), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    wg_fail(ctx, WG_ERR_CUDA, "wg_preview_plan_create", e);
    wg_preview_plan_destroy(pl);
    return WG_ERR_CUDA;
  }
  *out = pl;
  return WG_OK;
}

int wg_preview_plan_destroy(wg_preview_plan *pl)
{
  if (!pl) return WG_OK;
  wg_device_guard guard(pl->ctx->device);
  cudaStreamSynchronize(pl->ctx->stream);
  cudaFree(pl->d_offsets); cudaFree(pl->d_tiles); cudaFree(pl->d_fir);
  cudaFree(pl->d_zmp); cudaFree(pl->d_state); cudaFree(pl->d_com); cudaFree(pl->d_zmpout);
  delete pl;
  return WG_OK;
}

int64_t wg_preview_plan_total_steps(const wg_preview_plan *pl) { return pl ? pl->total_steps : 0; }
int64_t wg_preview_plan_total_samples(const wg_preview_plan *pl) { return pl ? pl->total_samples : 0; }

static int preview_launch(wg_ctx *ctx, wg_preview_plan *pl, const double *d_zmp, double *d_state,
                          double *d_com, double *d_zmpout, int simulation)
{
  if (pl->n_tiles == 0) return WG_OK;
  const int NLpad = (pl->NL + FIR_R - 1) / FIR_R * FIR_R;
  const int span = FIR_TILE + NLpad;
  const size_t smem = sizeof(double2) * (size_t)(span + (span >> 3) + 2);
  static bool attr_set = false;
  if (!attr_set) {
    WG_CUDA(ctx, cudaFuncSetAttribute(preview_fir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set = true;
  }
  if (smem > 64 * 1024) return wg_fail(ctx, WG_ERR_INVALID, "preview window too large for the FIR tile");
  wg_prof_start(ctx, WG_K_PREVIEW_FIR);
  preview_fir_kernel<<<pl->n_tiles, FIR_THREADS, smem, ctx->stream>>>(
      pl->d_tiles, pl->d_offsets, reinterpret_cast<const double2 *>(d_zmp), pl->d_fir);
  wg_prof_stop(ctx);
  WG_LAUNCHED(ctx);
  const int threads = 64, blocks = (2 * pl->B + threads - 1) / threads;
  wg_prof_start(ctx, WG_K_PREVIEW_RECUR);
  preview_recur_kernel<<<blocks, threads, 0, ctx->stream>>>(
      pl->B, pl->d_offsets, reinterpret_cast<const double2 *>(d_zmp), pl->d_fir, d_state, d_com, d_zmpout,
      simulation);
  wg_prof_stop(ctx);
  WG_LAUNCHED(ctx);
  return WG_OK;
}

int wg_preview_run_batch(wg_ctx *ctx, wg_preview_plan *pl, int mem, const double *zmpref_xy, double *state,
                         double *com_out, double *zmp_out, int simulation)
{
  if (!ctx || !pl || pl->ctx != ctx || !state || (!zmpref_xy && pl->total_samples > 0)) return WG_ERR_INVALID;
  if (!ctx->preview_ready || ctx->preview_gains.NL != pl->NL)
    return wg_fail(ctx, WG_ERR_NOT_READY, "gains changed since the plan was created");
  wg_device_guard guard(ctx->device);
  if (pl->B == 0) return WG_OK;
  if (mem == WG_MEM_DEVICE)
    return preview_launch(ctx, pl, zmpref_xy, state, com_out, zmp_out, simulation);
  if (mem != WG_MEM_HOST) return WG_ERR_INVALID;
  const size_t ns = (size_t)pl->total_samples;
  if (!pl->d_zmp) WG_CUDA(ctx, cudaMalloc(&pl->d_zmp, sizeof(double) * 2 * std::max<size_t>(1, ns)));
  if (!pl->d_state) WG_CUDA(ctx, cudaMalloc(&pl->d_state, sizeof(double) * 8 * pl->B));
  if (com_out && !pl->d_com) {  // zero once: rows past a trajectory's last step read back as 0 in host mode
    WG_CUDA(ctx, cudaMalloc(&pl->d_com, sizeof(double) * 6 * std::max<size_t>(1, ns)));
    WG_CUDA(ctx, cudaMemsetAsync(pl->d_com, 0, sizeof(double) * 6 * std::max<size_t>(1, ns), ctx->stream));
  }
  if (zmp_out && !pl->d_zmpout) {
    WG_CUDA(ctx, cudaMalloc(&pl->d_zmpout, sizeof(double) * 2 * std::max<size_t>(1, ns)));
    WG_CUDA(ctx, cudaMemsetAsync(pl->d_zmpout, 0, sizeof(double) * 2 * std::max<size_t>(1, ns), ctx->stream));
  }
  WG_CUDA(ctx, cudaMemcpyAsync(pl->d_zmp, zmpref_xy, sizeof(double) * 2 * ns, cudaMemcpyHostToDevice, ctx->stream));
  WG_CUDA(ctx, cudaMemcpyAsync(pl->d_state, state, sizeof(double) * 8 * pl->B, cudaMemcpyHostToDevice, ctx->stream));
  int rc = preview_launch(ctx, pl, pl->d_zmp, pl->d_state, com_out ? pl->d_com : nullptr,
                          zmp_out ? pl->d_zmpout : nullptr, simulation);
  if (rc != WG_OK) return rc;
  WG_CUDA(ctx, cudaMemcpyAsync(state, pl->d_state, sizeof(double) * 8 * pl->B, cudaMemcpyDeviceToHost, ctx->stream));
  if (com_out)
    WG_CUDA(ctx, cudaMemcpyAsync(com_out, pl->d_com, sizeof(double) * 6 * ns, cudaMemcpyDeviceToHost, ctx->stream));
  if (zmp_out)
    WG_CUDA(ctx, cudaMemcpyAsync(zmp_out, pl->d_zmpout, sizeof(double) * 2 * ns, cudaMemcpyDeviceToHost, ctx->stream));
  WG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return WG_OK;
}

int wg_preview_one_iteration(wg_ctx *ctx, double *x, double *y, double *sxzmp, double *syzmp,
                             const double *window_xy, int n_available, double *zmpx2, double *zmpy2,
                             int simulation)
{
  if (!ctx || !x || !y || !sxzmp || !syzmp || !window_xy || !zmpx2 || !zmpy2) return WG_ERR_INVALID;
  if (!ctx->preview_ready) return wg_fail(ctx, WG_ERR_NOT_READY, "wg_preview_set_gains not called");
  const int NL = ctx->preview_gains.NL;
  if (n_available < NL) return wg_fail(ctx, WG_ERR_WINDOW, "ZMPPositions.size()<m_SizeOfPreviewWindow");
  int64_t offs[2] = {0, NL};
  wg_preview_plan *pl = nullptr;
  int rc = wg_preview_plan_create(ctx, 1, offs, &pl);
  if (rc != WG_OK) return rc;
  double st[8] = {x[0], x[1], x[2], y[0], y[1], y[2], *sxzmp, *syzmp};
  std::vector<double> zo(2 * NL);
  rc = wg_preview_run_batch(ctx, pl, WG_MEM_HOST, window_xy, st, nullptr, zo.data(), simulation);
  wg_preview_plan_destroy(pl);
  if (rc != WG_OK) return rc;
  for (int i = 0; i < 3; ++i) { x[i] = st[i]; y[i] = st[3 + i]; }
  *sxzmp = st[6]; *syzmp = st[7];
  *zmpx2 = zo[0]; *zmpy2 = zo[1];
  return WG_OK;
}

}  // extern "C"
