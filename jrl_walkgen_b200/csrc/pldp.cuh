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
  double PuT[PLDP_N * PLDP_N];        // Pu transposed: PuT[i][k] = Pu[k][i] (lane-contiguous reads of one sample's column)
};

// SimilarConstraints of one problem + the scratch ComputeAlpha needs to honour them (see pldp_solve_warp); null = none
struct PldpSim {
  const int *similar;     // [m]
  double *t1;             // [128] per warp
  unsigned *amask;        // [4]   per warp: activity bit of every row
};

struct PldpWarp {
  double L[PLDP_KMAX * (PLDP_KMAX + 1) / 2];   // packed lower triangle, row i at i(i+1)/2
  double prev_zmp[PLDP_U];
  double vec[2][PLDP_U];                        // broadcast staging: c / d and Vk (one LDS instead of two SHFL per read)
  int active[PLDP_KMAX];
};

__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ---- the constraint matrix (m+1) x 32 as the solver sees it: M.row(r).at(c) ------------------------------------
// DenseMat: column-major array with leading dimension ld, as the reference receives it (global or shared memory).
struct DenseMat {
  static constexpr int kHalfUnroll = 2;   // the two axis halves are one flat column loop
  static constexpr int kColUnroll = 2;    // ComputeAlpha column loop: 2 x 4 rows in flight (64-register budget; 4 measured 5 % slower)
  const double *A;
  int ld;
  struct Row {
    const double *p; int ld;
    __device__ __forceinline__ double at(int c) const { return p[c * ld]; }   // 32-bit index: (m+1) * 32 < 2^31
    // the same element addressed as (axis half h, column kk of the half); cf = coef(h)
    __device__ __forceinline__ double coef(int) const { return 0.0; }
    __device__ __forceinline__ double at_h(double, int h, int kk) const { return p[(h * PLDP_N + kk) * ld]; }
  };
  __device__ __forceinline__ Row row(int r) const { return Row{A + r, ld}; }
};
// RankMat: the matrix BuildConstraintMatrices writes (ZMPConstrainedQPFastFormulation.cpp:885-905), never materialised:
// element (r, k + 16 ax) = a[ax][r] * Pu[k * 16 + i_r], the same single IEEE multiplication the reference stores.
struct RankMat {
  static constexpr int kHalfUnroll = 1;   // keep the loop over the axis halves rolled (code size)
  static constexpr int kColUnroll = 4;    // 128-register budget: 4 x 4 rows in flight (3 % faster than 2)
  const double *a0, *a1;   // [m]   A_r(0), A_r(1)
  const unsigned char *ri; // [m]   previewed sample of row r
  const double *Pu;        // [16][16] m_Pu
  const double *PuT;       // [16][16] its transpose
  struct Row {
    double a0, a1; const double *pu; const double *put;
    // one row read ACROSS lanes (lane = column c): the 16 entries of sample i_r are contiguous in PuT (reading them from
    // Pu[k][i_r] puts the 16 lanes of a half-warp on one shared-memory bank pair: a 16-way conflict per load)
    __device__ __forceinline__ double at(int c) const { return mul(c < PLDP_N ? a0 : a1, put[c & (PLDP_N - 1)]); }
    __device__ __forceinline__ double coef(int h) const { return h ? a1 : a0; }
    // a lane walking down ITS row (kk serial): lanes hold different samples i_r, contiguous in Pu[kk][.]
    __device__ __forceinline__ double at_h(double cf, int, int kk) const { return mul(cf, pu[kk * PLDP_N]); }
  };
  __device__ __forceinline__ Row row(int r) const { const int i = ri[r]; return Row{a0[r], a1[r], Pu + i, PuT + i * PLDP_N}; }
};

// OptCholesky::UpdateCholeskyMatrixFortran (OptCholesky.cpp:171-223): row `i` of L for the active rows act[0..i].
// Lane j computes M(i,j) = A_act[i] . A_act[j] in the reference's order and then the forward recurrence
// L(i,j) = (M(i,j) - sum_{k<j} L(i,k) L(j,k)) / L(j,j), with the L(i,k) broadcast as they become final.
template <class Mat>
__device__ __noinline__ void chol_add_row(double *L, const int *act, int i, const Mat &M, int lane)
{
  double r = 0.0;
  {
    const typename Mat::Row ri = M.row(act[i]), rj = M.row(act[lane <= i ? lane : i]);
#pragma unroll(Mat::kHalfUnroll)
    for (int h = 0; h < 2; ++h) {
      const double ci = ri.coef(h), cj = rj.coef(h);
#pragma unroll 4
      for (int kk = 0; kk < PLDP_N; ++kk) r = add(r, mul(ri.at_h(ci, h, kk), rj.at_h(cj, h, kk)));
    }
  }
  double lij = 0.0;
#pragma unroll 1
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

// Rows 0..k-1 of L at once (the hot start re-activates the kept constraints before the first iteration: the reference
// calls AddActiveConstraint k times, PLDPSolver.cpp:763-778).  Lane i owns row i; column by column, every entry is the same
// expression in the same order as in chol_add_row - L(i,j) = (M(i,j) - sum_{t<j} L(i,t) L(j,t)) / L(j,j) with t ascending -
// so the factor is bit-identical, but the k(k+1)/2 serial divisions become k steps of one parallel division.
template <class Mat>
__device__ __noinline__ void chol_build(double *L, const int *act, int k, const Mat &M, int lane)
{
  const bool own = lane < k;
  const typename Mat::Row ri = M.row(act[own ? lane : 0]);
#pragma unroll 1
  for (int j = 0; j < k; ++j) {
    const typename Mat::Row rj = M.row(act[j]);
    double r = 0.0;
#pragma unroll(Mat::kHalfUnroll)
    for (int h = 0; h < 2; ++h) {
      const double ci = ri.coef(h), cj = rj.coef(h);
#pragma unroll 4
      for (int kk = 0; kk < PLDP_N; ++kk) r = add(r, mul(ri.at_h(ci, h, kk), rj.at_h(cj, h, kk)));
    }
    if (own && lane >= j) {
      const double *Li = L + tri(lane), *Lj = L + tri(j);
#pragma unroll 2
      for (int t = 0; t < j; ++t) r = add(r, -mul(Li[t], Lj[t]));
    }
    double d = 0.0;
    if (lane == j) { d = sqrt(r); L[tri(j) + j] = d; }
    d = bcast(d, j);
    if (own && lane > j) L[tri(lane) + j] = r / d;
    __syncwarp();
  }
}

// PLDPSolver::SolveProblem for the problem (M: m x 32 constraint matrix, bv, Dl = D[lane], zr, xk) by the calling warp.
// use_prev: start from the shifted previous ZMP solution prev_zmp[32] (hot start, not the first call); n_prev /
// prev_active / nr: constraints kept from the previous solve and NumberOfRemovedConstraints.
// Returns Vk (lane = entry); r.v2 / r.kproj: multipliers of the last projection (lane i < kproj), w.active[0..r.k).
struct PldpRes { int status, it, k, kproj; double v2; };

template <class Mat>
__device__ __forceinline__ double pldp_solve_warp(const PldpConsts &C, PldpWarp &w, const Mat &M, int m,
                                                  const double *bv, double Dl, const double *zr, const double *xk,
                                                  bool use_prev, const double *prev_zmp, int n_prev,
                                                  const int *prev_active, int nr, int max_iter, double tol, int lane,
                                                  PldpRes &r, const PldpSim *sim = nullptr)
{
  constexpr int N = PLDP_N;
  constexpr int SLABS = 4;                       // rows lane, lane+32, lane+64, lane+96: m <= 128
  int status = 0;
  // SimilarConstraints (sim): ComputeAlpha's `A_j = -A_i` reuse (PLDPSolver.cpp:570-590) takes tmp1[li] = -tmp1[li + s]
  // when s = similar[li] != 0 and row li + s was visited and is not active.  For flags that match the matrix (the only
  // kind the reference builds: the flagged row is the exact negation of the other) this is bit-identical to the direct
  // product, which is why the Dimitrov loop passes no flags; the stand-alone solver honours caller-supplied flags exactly
  // (chains included), and refuses flags that point forward or out of range (status 6: the reference would read its flag
  // array as an earlier call left it).  The second reuse (:600-610) can never fire and is not restated.

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
#pragma unroll 2
      for (int j = 0; j < N - 1; ++j) Vk = add(Vk, mul(C.iPu[j * N + ii], w.prev_zmp[j + 1 + N * ax]));
      Vk = add(Vk, mul(C.iPu[(N - 1) * N + ii], zr[N - 1 + N * ax]));
    } else {
#pragma unroll 2
      for (int j = 0; j < N; ++j) Vk = add(Vk, mul(C.iPu[j * N + ii], zr[j + N * ax]));
    }
  }
  // ---- hot start: re-activate the constraints kept from the previous solve (PLDPSolver.cpp:763-778)
  int k = 0;
  if (n_prev > 0) {
    const int np = n_prev;
#pragma unroll 1
    for (int i = 0; i < np && k < PLDP_KMAX; ++i) {
      const int idx = prev_active[i] - nr;
      if (idx >= 0 && idx < m) {
        if (lane == 0) w.active[k] = idx;
        ++k;
      }
    }
    __syncwarp();
    if (k > 0) chol_build(w.L, w.active, k, M, lane);
  }
  // activity flags of the rows this lane owns (rows lane, lane+32, lane+64, lane+96)
  unsigned mine = 0;
#pragma unroll 1
  for (int i = 0; i < k; ++i) { const int r = w.active[i]; if ((r & 31) == lane) mine |= 1u << (r >> 5); }
  int kproj = 0;         // size of the active set at the last projection (v2 is defined for lanes < kproj)
  const int ns = (m + 31) >> 5;   // slabs of 32 rows in use

  double v2 = 0.0;       // lane i < k holds v2[i] of the last projection
  int it = 0;
  bool cont = true;
  while (cont) {
    // ---- step 1: c = -D - Vk (PLDPSolver.cpp:805-807)
    const double c = add(-Dl, -Vk);
    __syncwarp();
    w.vec[0][lane] = c;
    w.vec[1][lane] = Vk;
    __syncwarp();
    // ---- step 2: ComputeProjectedDescentDirection (PLDPSolver.cpp:404-532)
    // v1 = E c : lane li owns active row li
    double v1 = 0.0;
    if (k > 0) {
      const typename Mat::Row row = M.row(w.active[lane < k ? lane : 0]);
#pragma unroll(Mat::kHalfUnroll)
      for (int h = 0; h < 2; ++h) {
        const double cf = row.coef(h);
#pragma unroll 4
        for (int kk = 0; kk < PLDP_N; ++kk) v1 = add(v1, mul(row.at_h(cf, h, kk), w.vec[0][h * PLDP_N + kk]));
      }
    }
    // forward substitution L y = v1 (:342-365): y[i] += -L(i,k) y[k] in k order, then / L(i,i) (skipped when 0)
    double y = v1;
#pragma unroll 1
    for (int i = 0; i < k; ++i) {
      if (lane == i) { const double dg = w.L[tri(i) + i]; if (dg != 0.0) y = y / dg; }
      const double yi = bcast(y, i);
      if (lane > i && lane < k) y = add(y, -mul(w.L[tri(lane) + i], yi));
    }
    // backward substitution L^T v2 = y (:367-400): v2[i] = (y[i] - sum_{k'=i+1}^{k-1} L(k',i) v2[k']) / L(i,i), the
    // sum taken in INCREASING k' as the reference does.  Lane k' forms its product in parallel; only the ordered
    // additions are serial (the products and the shuffles are off the dependency chain).
    v2 = y;
#pragma unroll 1
    for (int i = k - 1; i >= 0; --i) {
      const double p = (lane > i && lane < k) ? mul(w.L[tri(lane) + i], v2) : 0.0;
      double acc = bcast(v2, i);
#pragma unroll 2
      for (int kk = i + 1; kk < k; ++kk) acc = add(acc, -bcast(p, kk));
      acc = acc / w.L[tri(i) + i];
      if (lane == i) v2 = acc;
    }
    kproj = k;
    // d = c - E^T v2 (:509-518): lane li, sequential over the active rows
    double d = c;
#pragma unroll 1
    for (int j = 0; j < k; ++j) {
      const double vj = bcast(v2, j);
      d = add(d, -mul(M.row(w.active[j]).at(lane), vj));
    }
    __syncwarp();
    w.vec[0][lane] = d;
    __syncwarp();
    // ---- step 3: ComputeAlpha (:534-653): rows lane, lane+32, ... ; running minimum in row order.  The column index
    // runs in the OUTER loop so that the (up to 4) rows of a lane advance as independent dependency chains; each row's
    // own sum keeps the reference's order.
    double alpha = 10000000.0;
    int cand = -1;
    {
      double t1[SLABS];
      bool ok[SLABS];
#pragma unroll
      for (int s = 0; s < SLABS; ++s) { const int li = lane + 32 * s; ok[s] = (li < m) && !((mine >> s) & 1u); t1[s] = 0.0; }
      {
        typename Mat::Row rows[SLABS];
#pragma unroll
        for (int s = 0; s < SLABS; ++s) rows[s] = M.row((lane + 32 * s) < m ? lane + 32 * s : 0);
#pragma unroll(Mat::kHalfUnroll)
        for (int h = 0; h < 2; ++h) {
          double cf[SLABS];
#pragma unroll
          for (int s = 0; s < SLABS; ++s) cf[s] = rows[s].coef(h);
#pragma unroll(Mat::kColUnroll)
          for (int kk = 0; kk < PLDP_N; ++kk) {
            const double dj = w.vec[0][h * PLDP_N + kk];
#pragma unroll
            for (int s = 0; s < SLABS; ++s)
              if (s < ns) t1[s] = add(t1[s], mul(rows[s].at_h(cf[s], h, kk), dj));
          }
        }
        if (sim) {
          __syncwarp();
#pragma unroll
          for (int s = 0; s < SLABS; ++s) {
            if (lane + 32 * s < m) sim->t1[lane + 32 * s] = t1[s];
            const unsigned bits = __ballot_sync(0xffffffffu, (mine >> s) & 1u);
            if (lane == 0) sim->amask[s] = bits;
          }
          __syncwarp();
#pragma unroll
          for (int s = 0; s < SLABS; ++s) {
            if (!ok[s]) continue;
            const int li = lane + 32 * s;
            int rr = li;
            bool neg = false;
#pragma unroll 1
            for (;;) {
              const int sm = sim->similar[rr];
              if (sm == 0) break;
              const int t = rr + sm;
              if (t < 0 || t >= rr) { status = 6; break; }
              if ((sim->amask[t >> 5] >> (t & 31)) & 1u) break;      // active rows are skipped before their product
              rr = t; neg = !neg;
            }
            if (rr != li) t1[s] = neg ? -sim->t1[rr] : sim->t1[rr];
          }
        }
        bool want[SLABS];
        bool any = false;
#pragma unroll
        for (int s = 0; s < SLABS; ++s) { want[s] = ok[s] && t1[s] < 0.0; any = any || want[s]; }
        double best = 10000000.0; int besti = 0x7fffffff;
        if (__any_sync(0xffffffffu, any)) {
          double t2[SLABS];
          // slabs in which no row moves towards its bound (t1 < 0) need no second product: skip them as a whole
          bool wslab[SLABS];
#pragma unroll
          for (int s = 0; s < SLABS; ++s) wslab[s] = __any_sync(0xffffffffu, want[s]);
#pragma unroll
          for (int s = 0; s < SLABS; ++s) t2[s] = -bv[(lane + 32 * s) < m ? lane + 32 * s : 0];
#pragma unroll(Mat::kHalfUnroll)
          for (int h = 0; h < 2; ++h) {
            double cf[SLABS];
#pragma unroll
            for (int s = 0; s < SLABS; ++s) cf[s] = rows[s].coef(h);
#pragma unroll(Mat::kColUnroll)
            for (int kk = 0; kk < PLDP_N; ++kk) {
              const double vj = w.vec[1][h * PLDP_N + kk];
#pragma unroll
              for (int s = 0; s < SLABS; ++s)
                if (wslab[s]) t2[s] = add(t2[s], -mul(rows[s].at_h(cf[s], h, kk), vj));
            }
          }
#pragma unroll
          for (int s = 0; s < SLABS; ++s) {
            if (want[s]) {
              double t = t2[s];
              if (t > tol) status = 1;                // "PB ON constraint": the start point violates row li
              else if (t > 0.0) t = -tol;
              const double la = t / t1[s];
              if (la < best) { best = la; besti = lane + 32 * s; }
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
    }
    status = __reduce_max_sync(0xffffffffu, status);
    if (status == 6) break;
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
        chol_add_row(w.L, w.active, k, M, lane);
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
