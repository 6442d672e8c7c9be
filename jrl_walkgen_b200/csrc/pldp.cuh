// pldp.cuh - device side of the Dimitrov PLDP solver (one warp per problem), shared by pldp.cu (batched
// SolveProblem) and dimitrov.cu (the receding-horizon loop).  See pldp.cu for the design notes.
#pragma once
#include "wg_common.h"

namespace {

constexpr int PLDP_N = 16;            // m_CardV
constexpr int PLDP_U = 2 * PLDP_N;    // 32 = warp size
constexpr int PLDP_KMAX = 32;         // active-set capacity (= number of columns of E)
constexpr int PLDP_WARPS = 4;

struct PldpConsts {
  double iPu[PLDP_N * PLDP_N];        // row-major as handed to the PLDPSolver ctor
  double Px[PLDP_N * 3];
  double Pu[PLDP_N * PLDP_N];
  double iPuPx[PLDP_U * 6];           // PLDPSolver::PrecomputeiPuPx (PLDPSolver.cpp:264-285)
};

struct PldpWarp {
  double L[PLDP_KMAX * (PLDP_KMAX + 1) / 2];   // packed lower triangle, row i at i(i+1)/2
  double prev_zmp[PLDP_U];
  int active[PLDP_KMAX];
};

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// OptCholesky::UpdateCholeskyMatrixFortran (OptCholesky.cpp:171-223): row `i` of L for the active rows act[0..i].
// Lane j computes M(i,j) = A_act[i] . A_act[j] in the reference's order and then the forward recurrence
// L(i,j) = (M(i,j) - sum_{k<j} L(i,k) L(j,k)) / L(j,j), with the L(i,k) broadcast as they become final.
__device__ void chol_add_row(double *L, const int *act, int i, const double *A, int ld, int lane)
{
  double r = 0.0;
  if (lane <= i) {
    const double *ri = A + act[i], *rj = A + act[lane];
#pragma unroll 4
    for (int k = 0; k < PLDP_U; ++k) r = add(r, mul(ri[(size_t)k * ld], rj[(size_t)k * ld]));
  }
  double lij = 0.0;
  for (int j = 0; j <= i; ++j) {
    // lane j finalises L(i,j)
    if (lane == j) lij = (j != i) ? r / L[tri(j) + j] : sqrt(r);
    const double v = bcast(lij, j);
    // for the diagonal (lane == i) the second factor is the new row itself, i.e. the value just finalised
    if (lane > j && lane <= i) r = add(r, -mul(v, (lane == i) ? v : L[tri(lane) + j]));
  }
  if (lane <= i) L[tri(i) + lane] = lij;
  __syncwarp();
}

// PLDPSolver::SolveProblem for the problem (A: (m+1) x 32 column-major with leading dimension ld, bv, Dl = D[lane], zr,
// xk) by the calling warp.  use_prev: start from the shifted previous ZMP solution prev_zmp[32] (hot start, not the
// first call); n_prev / prev_active / nr: constraints kept from the previous solve and NumberOfRemovedConstraints.
// Returns Vk (lane = entry); r.v2 / r.kproj: multipliers of the last projection (lane i < kproj), w.active[0..r.k).
struct PldpRes { int status, it, k, kproj; double v2; };

__device__ __forceinline__ double pldp_solve_warp(const PldpConsts &C, PldpWarp &w, const double *A, int ld, int m,
                                                  const double *bv, double Dl, const double *zr, const double *xk,
                                                  bool use_prev, const double *prev_zmp, int n_prev,
                                                  const int *prev_active, int nr, int max_iter, double tol, int lane,
                                                  PldpRes &r)
{
  constexpr int N = PLDP_N;
  int status = 0;
  // `similar` is accepted for interface parity only: the A_j = -A_i reuse (PLDPSolver.cpp:570-590) yields products
  // bit-identical to computing every row directly, which is what the lanes do

  // ---- ComputeInitialSolution (PLDPSolver.cpp:287-340): lane = i (x part) or i + N (y part)
  const int ii = lane & (N - 1), ax = lane >> 4;
  if (use_prev) w.prev_zmp[lane] = prev_zmp[lane];
  __syncwarp();
  double Vk = 0.0;
  {
    const double *ipx = C.iPuPx + lane * 6 + 3 * ax;
#pragma unroll
    for (int j = 0; j < 3; ++j) Vk = add(Vk, -mul(ipx[j], xk[3 * ax + j]));
    if (use_prev) {
      for (int j = 0; j < N - 1; ++j) Vk = add(Vk, mul(C.iPu[j * N + ii], w.prev_zmp[j + 1 + N * ax]));
      Vk = add(Vk, mul(C.iPu[(N - 1) * N + ii], zr[N - 1 + N * ax]));
    } else {
      for (int j = 0; j < N; ++j) Vk = add(Vk, mul(C.iPu[j * N + ii], zr[j + N * ax]));
    }
  }
  // ---- hot start: re-activate the constraints kept from the previous solve (PLDPSolver.cpp:763-778)
  int k = 0;
  if (n_prev > 0) {
    const int np = n_prev;
    for (int i = 0; i < np && k < PLDP_KMAX; ++i) {
      const int idx = prev_active[i] - nr;
      if (idx >= 0 && idx < m) {
        if (lane == 0) w.active[k] = idx;
        __syncwarp();
        chol_add_row(w.L, w.active, k, A, ld, lane);
        ++k;
      }
    }
  }
  // activity flags of the rows this lane owns (rows lane, lane+32, lane+64, lane+96)
  unsigned mine = 0;
  for (int i = 0; i < k; ++i) { const int r = w.active[i]; if ((r & 31) == lane) mine |= 1u << (r >> 5); }
  int kproj = 0;         // size of the active set at the last projection (v2 is defined for lanes < kproj)

  double v2 = 0.0;       // lane i < k holds v2[i] of the last projection
  int it = 0;
  bool cont = true;
  while (cont) {
    // ---- step 1: c = -D - Vk (PLDPSolver.cpp:805-807)
    const double c = add(-Dl, -Vk);
    // ---- step 2: ComputeProjectedDescentDirection (PLDPSolver.cpp:404-532)
    // v1 = E c : lane li owns active row li
    double v1 = 0.0;
    {
      const double *row = A + ((lane < k) ? w.active[lane] : 0);
#pragma unroll 4
      for (int j = 0; j < PLDP_U; ++j) {
        const double cj = bcast(c, j);
        if (lane < k) v1 = add(v1, mul(row[(size_t)j * ld], cj));
      }
    }
    // forward substitution L y = v1 (:342-365): y[i] += -L(i,k) y[k] in k order, then / L(i,i) (skipped when 0)
    double y = v1;
    for (int i = 0; i < k; ++i) {
      if (lane == i) { const double dg = w.L[tri(i) + i]; if (dg != 0.0) y = y / dg; }
      const double yi = bcast(y, i);
      if (lane > i && lane < k) y = add(y, -mul(w.L[tri(lane) + i], yi));
    }
    // backward substitution L^T v2 = y (:367-400): v2[i] = (y[i] - sum_{k'=i+1}^{k-1} L(k',i) v2[k']) / L(i,i), the
    // sum taken in INCREASING k' as the reference does.  Lane k' forms its product in parallel; only the ordered
    // additions are serial (the products and the shuffles are off the dependency chain).
    v2 = y;
    for (int i = k - 1; i >= 0; --i) {
      const double p = (lane > i && lane < k) ? mul(w.L[tri(lane) + i], v2) : 0.0;
      double acc = bcast(v2, i);
      for (int kk = i + 1; kk < k; ++kk) acc = add(acc, -bcast(p, kk));
      acc = acc / w.L[tri(i) + i];
      if (lane == i) v2 = acc;
    }
    kproj = k;
    // d = c - E^T v2 (:509-518): lane li, sequential over the active rows
    double d = c;
    for (int j = 0; j < k; ++j) {
      const double vj = bcast(v2, j);
      d = add(d, -mul(A[w.active[j] + (size_t)lane * ld], vj));
    }
    // ---- step 3: ComputeAlpha (:534-653): rows lane, lane+32, ... ; running minimum in row order
    double alpha = 10000000.0;
    int cand = -1;
    {
      double best = 10000000.0; int besti = 0x7fffffff;
      for (int s = 0; s * 32 < m; ++s) {
        const int li = lane + 32 * s;
        double t1 = 0.0, t2 = 0.0;
        const bool mineok = (li < m) && !((mine >> s) & 1u);
        const double *row = A + (li < m ? li : 0);
#pragma unroll 4
        for (int j = 0; j < PLDP_U; ++j) {
          const double dj = bcast(d, j);
          if (mineok) t1 = add(t1, mul(row[(size_t)j * ld], dj));
        }
        const unsigned neg = __ballot_sync(0xffffffffu, mineok && t1 < 0.0);
        if (neg) {
          t2 = mineok ? -bv[li < m ? li : 0] : 0.0;
#pragma unroll 4
          for (int j = 0; j < PLDP_U; ++j) {
            const double vj = bcast(Vk, j);
            if (mineok && t1 < 0.0) t2 = add(t2, -mul(row[(size_t)j * ld], vj));
          }
          if (mineok && t1 < 0.0) {
            if (t2 > tol) status = 1;                // "PB ON constraint": the start point violates row li
            else if (t2 > 0.0) t2 = -tol;
            const double la = t2 / t1;
            if (la < best) { best = la; besti = li; }
          }
        }
      }
      // the reference keeps the FIRST row (in index order) that attains the minimum
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
      }
      if (best < alpha) { alpha = best; if (alpha < 1.0) cand = besti; }
    }
    status = __reduce_max_sync(0xffffffffu, status);
    if (alpha >= 1.0) { alpha = 1.0; cont = false; }
    if (alpha < 0.0) { status = 2; cont = false; }     // the reference calls exit(0) here (:822-828)
    // ---- new solution (:830-834)
    if (status != 2) Vk = add(Vk, mul(alpha, d));
    if (cont) {
      if (k >= PLDP_KMAX || cand < 0) { status = 3; cont = false; }
      else {
        if (lane == 0) w.active[k] = cand;
        if ((cand & 31) == lane) mine |= 1u << (cand >> 5);
        __syncwarp();
        chol_add_row(w.L, w.active, k, A, ld, lane);
        ++k;
      }
    }
    ++it;
    if (it >= max_iter && cont) { cont = false; status = status ? status : 4; }   // stands in for the 1.3 ms cap
  }
  r.status = status; r.it = it; r.k = k; r.kproj = kproj; r.v2 = v2;
  return Vk;
}

}  // namespace
