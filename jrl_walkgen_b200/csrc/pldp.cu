// pldp.cu - batched Dimitrov PLDP solver and OptCholesky for sm_100a.
//
// Replaces PLDPSolver::SolveProblem and its helpers (src/Mathematics/PLDPSolver.cpp:287-1007) and
// OptCholesky (src/Mathematics/OptCholesky.cpp:92-302) for whole batches of independent problems.
//
// One warp owns one problem.  The control vector has 2N = 32 entries = one per lane, so the iterate V_k, the
// descent direction c, the projected direction d live in registers (one double per lane) and move with shuffles;
// the Cholesky factor of E E^T (at most 32 x 32: PLDP never drops a constraint and E has 32 columns) is packed in
// shared memory.  The constraint matrix (m+1) x 32, column-major exactly as the reference receives it, is read
// from global memory / L2 by "row per lane" loops (consecutive lanes read consecutive rows: coalesced).
//
// Arithmetic follows the reference statement by statement with NON-fused multiplies and adds (__dmul_rn/__dadd_rn)
// in the reference's summation order throughout, so that on identical inputs the iterates, the step lengths and therefore the
// sequence of activated constraints are the same as the reference's x86-64 object code (which has no FMA
// contraction at -O3 without -march).  Differences by design: the wall-clock cap of 1.3 ms (PLDPSolver.cpp:69,
// :890-900) becomes an iteration cap, and the paths on which the reference calls exit(0) (:822-828) or prints
// return a status instead.
#include "wg_common.h"
#include <algorithm>
#include <vector>
#include <cmath>

#include "pldp.cuh"

namespace {

struct PldpHost {
  PldpConsts h;
  PldpConsts *d = nullptr;
  bool ready = false;
  // staging for WG_MEM_HOST calls
  void *buf[10] = {nullptr};
  size_t cap[10] = {0};
  int *d_next = nullptr;   // work counter of pldp_kernel
  // WG_MEM_HOST pipeline
  cudaStream_t up = nullptr, down = nullptr;
  cudaEvent_t ev_up[4] = {nullptr}, ev_k[4] = {nullptr}, ev0 = nullptr;
};

// One instance per warp.
__global__ void __launch_bounds__(PLDP_WARPS * 32, 8)
pldp_kernel(int B, const PldpConsts *__restrict__ Cp, const double *__restrict__ D, const int *__restrict__ mvec,
            const double *__restrict__ DPu, long long dpu_stride, const double *__restrict__ DPx, long long dpx_stride,
            const double *__restrict__ ZMPRef, const double *__restrict__ XkYk, double *__restrict__ X,
            const int *__restrict__ similar, long long similar_stride, const int *__restrict__ nremoved,
            const int *__restrict__ starting, wg_pldp_state *__restrict__ hot, int hot_start, int max_iter,
            double tol, wg_pldp_info *__restrict__ info, int a_cap, int *__restrict__ next_problem)
{
  __shared__ PldpWarp ws[PLDP_WARPS];
  extern __shared__ __align__(16) double sA[];   // a_cap doubles per warp: the instance's constraint matrix (0: read it from L2)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const PldpConsts &C = *Cp;
  PldpWarp &w = ws[warp];
  constexpr int N = PLDP_N;
  // iteration counts differ (5-33 on the bench workload): every warp takes its next problem from a work counter
  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(next_problem, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= B) break;
    const int m = mvec[b];
    const int ld = m + 1;
    const double *A = DPu + (size_t)b * dpu_stride;
    if (a_cap > 0 && ld * PLDP_U <= a_cap) {
      // stage the (m+1) x 32 column-major matrix once: every iteration re-reads all of it
      double *dst = sA + (size_t)warp * a_cap;
      const int n = ld * PLDP_U;
      __syncwarp();
      if ((((size_t)A) & 15) == 0) {
        const double2 *src2 = reinterpret_cast<const double2 *>(A);
        double2 *dst2 = reinterpret_cast<double2 *>(dst);
        for (int e = lane; e < (n >> 1); e += 32) dst2[e] = __ldg(src2 + e);
        if ((n & 1) && lane == 0) dst[n - 1] = A[n - 1];
      } else {
        for (int e = lane; e < n; e += 32) dst[e] = A[e];
      }
      __syncwarp();
      A = dst;
    }
    const double *bv = DPx + (size_t)b * dpx_stride;
    const bool start = starting ? starting[b] != 0 : true;
    const double Dl = D[(size_t)b * PLDP_U + lane];
    const double *zr = ZMPRef + (size_t)b * PLDP_U;
    const double *xk = XkYk + (size_t)b * 6;
    const bool hs = hot && hot_start;
    PldpRes r;
    const DenseMat M{A, ld};
    const double Vk = pldp_solve_warp(C, w, M, m, bv, Dl, zr, xk, hs && !start, hs ? hot[b].prev_zmp : nullptr,
                                      hs ? hot[b].n_prev : 0, hs ? hot[b].prev_active : nullptr,
                                      nremoved ? nremoved[b] : 0, max_iter, tol, lane, r);
    const int status = r.status, it = r.it, k = r.k, kproj = r.kproj;
    const double v2 = r.v2;
    const int ii = lane & (N - 1), ax = lane >> 4;
    // ---- results
    X[(size_t)b * PLDP_U + lane] = Vk;
    const double x0 = bcast(Vk, 0), xn = bcast(Vk, N);
    int rc = 0;
    if (isnan(x0) || isnan(xn) || isinf(x0) || isinf(xn)) rc = -1;   // PLDPSolver.cpp:955-964
    if (hot && hot_start) {
      // keep the rows whose multiplier is negative (:909-920) and store the ZMP solution (:1009-1032)
      const unsigned keep = __ballot_sync(0xffffffffu, lane < kproj && v2 < 0.0);
      if (lane < kproj && v2 < 0.0) hot[b].prev_active[__popc(keep & ((1u << lane) - 1u))] = w.active[lane];
      if (lane == 0) hot[b].n_prev = __popc(keep);
      double z = 0.0;
#pragma unroll 2
      for (int j = 0; j < N; ++j) z = add(z, mul(C.Pu[j * N + ii], bcast(Vk, j + N * ax)));
#pragma unroll
      for (int j = 0; j < 3; ++j) z = add(z, mul(C.Px[ii * 3 + j], xk[3 * ax + j]));
      hot[b].prev_zmp[lane] = z;
    }
    if (info) {
      if (lane == 0) { info[b].rc = rc; info[b].status = status; info[b].iterations = it; info[b].n_active = k; }
      info[b].active[lane] = (lane < k) ? w.active[lane] : -1;
    }
    __syncwarp();
  }
}

// OptCholesky batched: rows k0..k-1 of L for the active rows `rows` of each instance's A.
//   mode 1 (MODE_FORTRAN): A column-major, element (r, c) at A[r + c (nb_constraints + 1)]  (OptCholesky.cpp:171-223)
//   mode 0 (MODE_NORMAL):  A row-major,    element (r, c) at A[r card_u + c]                 (OptCholesky.cpp:123-169)
// L is row-major with leading dimension nb_max (the caller-owned storage the reference writes into).
__global__ void __launch_bounds__(128)
optchol_kernel(int B, int mode, int nb_max, int card_u, int nb_constraints, const double *__restrict__ A,
               long long a_stride, const int *__restrict__ rows, int rows_stride, int k0, int k1,
               double *__restrict__ Lout, long long l_stride)
{
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x * 4 + warp; b < B; b += gridDim.x * 4) {
    const double *Ab = A + (size_t)b * a_stride;
    const int *act = rows + (size_t)b * rows_stride;
    double *L = Lout + (size_t)b * l_stride;
    const long long rs = mode ? 1 : card_u, cs = mode ? (nb_constraints + 1) : 1;
    for (int i = k0; i < k1; ++i) {
      const double *ri = Ab + (size_t)act[i] * rs;
      // columns j = 0..i, 32 at a time; the recurrence over j is sequential, lane (j % 32) finalises L(i,j)
      for (int j = 0; j <= i; ++j) {
        const double *rj = Ab + (size_t)act[j] * rs;
        // M(i,j) in the reference's order, then r -= L(i,k) L(j,k) for k < j: one lane does the serial sum
        // (bit-faithful); different (i,j) pairs cannot run ahead because of the recurrence on L(i,k<j).
        if (lane == 0) {
          double r = 0.0;
          for (int c = 0; c < card_u; ++c) r = add(r, mul(ri[c * cs], rj[c * cs]));
          const double *Li = L + (size_t)i * nb_max, *Lj = L + (size_t)j * nb_max;
          for (int kk = 0; kk < j; ++kk) r = add(r, -mul(Li[kk], Lj[kk]));
          L[(size_t)i * nb_max + j] = (j != i) ? r / Lj[j] : sqrt(r);
        }
        __syncwarp();
      }
    }
  }
}

// ComputeNormalCholeskyOnANormal (OptCholesky.cpp:225-259) + ComputeInverseCholeskyNormal (:261-302), one warp each.
__global__ void __launch_bounds__(128)
optchol_full_kernel(int B, int n, const double *__restrict__ A, double *__restrict__ Lout, double *__restrict__ iLout,
                    int inv_size)
{
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x * 4 + warp; b < B; b += gridDim.x * 4) {
    const double *Ab = A ? A + (size_t)b * n * n : nullptr;
    double *L = Lout + (size_t)b * n * n;
    double *iL = iLout ? iLout + (size_t)b * n * n : nullptr;
    if (Ab) {
      for (int i = 0; i < n; ++i) {
        for (int j = 0; j <= i; ++j) {
          if (lane == 0) {
            double r = Ab[(size_t)i * n + j];
            for (int kk = 0; kk < j; ++kk) r = add(r, -mul(L[(size_t)i * n + kk], L[(size_t)j * n + kk]));
            L[(size_t)i * n + j] = (j != i) ? r / L[(size_t)j * n + j] : sqrt(r);
          }
          __syncwarp();
        }
      }
    }
    if (iL) {
      // columns lj are independent of each other: lane-strided over lj, serial down the column as the reference
      for (int lj = inv_size - 1 - lane; lj >= 0; lj -= 32) {
        const double inv = 1 / L[(size_t)lj * n + lj];
        iL[(size_t)lj * n + lj] = inv;
      }
      __syncwarp();
      // iL(li,lj) = -iL(lj,lj) * sum_{lk=lj+1..} iL(li,lk) L(lk,lj): needs iL(li, lk>lj) -> process lj descending
      for (int lj = inv_size - 1; lj >= 0; --lj) {
        const double inv = iL[(size_t)lj * n + lj];
        for (int li = lj + 1 + lane; li < inv_size; li += 32) {
          double r = 0.0;
          for (int lk = lj + 1; lk < inv_size; ++lk) r = add(r, mul(iL[(size_t)li * n + lk], L[(size_t)lk * n + lj]));
          iL[(size_t)li * n + lj] = mul(-inv, r);
        }
        __syncwarp();
      }
    }
  }
}

PldpHost *pldp_of(wg_ctx *ctx)
{
  if (!ctx->pldp) ctx->pldp = new PldpHost();
  return static_cast<PldpHost *>(ctx->pldp);
}

int ensure(wg_ctx *ctx, PldpHost *p, int slot, size_t bytes)
{
  if (p->cap[slot] >= bytes) return WG_OK;
  WG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(p->buf[slot]);
  p->buf[slot] = nullptr; p->cap[slot] = 0;
  WG_CUDA(ctx, cudaMalloc(&p->buf[slot], bytes ? bytes : 8));
  p->cap[slot] = bytes;
  return WG_OK;
}

int grid_for(wg_ctx *ctx, int B, int warps) { int g = (B + warps - 1) / warps; int cap = ctx->sm_count * 8; return g < cap ? (g > 0 ? g : 1) : cap; }

}  // namespace

extern "C" const void *wgi_pldp_device_consts(wg_ctx *ctx)
{
  PldpHost *p = static_cast<PldpHost *>(ctx->pldp);
  return (p && p->ready) ? p->d : nullptr;
}

void wg_pldp_release(wg_ctx *ctx)
{
  if (!ctx->pldp) return;
  PldpHost *p = static_cast<PldpHost *>(ctx->pldp);
  cudaFree(p->d); cudaFree(p->d_next);
  if (p->up) cudaStreamDestroy(p->up);
  if (p->down) cudaStreamDestroy(p->down);
  for (int c = 0; c < 4; ++c) { if (p->ev_up[c]) cudaEventDestroy(p->ev_up[c]); if (p->ev_k[c]) cudaEventDestroy(p->ev_k[c]); }
  if (p->ev0) cudaEventDestroy(p->ev0);
  for (void *b : p->buf) cudaFree(b);
  delete p;
  ctx->pldp = nullptr;
}

extern "C" {

int wg_pldp_set_constants(wg_ctx *ctx, int card_u, const double *iPu, const double *Px, const double *Pu)
{
  if (!ctx || !iPu || !Px || !Pu) return WG_ERR_INVALID;
  if (card_u != PLDP_N) return wg_fail(ctx, WG_ERR_INVALID, "PLDP kernels are built for CardU = 16 (QP_N of the reference)");
  wg_device_guard guard(ctx->device);
  PldpHost *p = pldp_of(ctx);
  const int N = PLDP_N;
  std::memcpy(p->h.iPu, iPu, sizeof p->h.iPu);
  std::memcpy(p->h.Px, Px, sizeof p->h.Px);
  std::memcpy(p->h.Pu, Pu, sizeof p->h.Pu);
  // PLDPSolver::PrecomputeiPuPx, PLDPSolver.cpp:264-285 (same summation order)
  std::memset(p->h.iPuPx, 0, sizeof p->h.iPuPx);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < N; ++k) {
        const double tmp = iPu[k * N + i] * Px[k * 3 + j];
        p->h.iPuPx[i * 6 + j] += tmp;
        p->h.iPuPx[(i + N) * 6 + j + 3] += tmp;
      }
  if (!p->d) WG_CUDA(ctx, cudaMalloc(&p->d, sizeof(PldpConsts)));
  WG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  WG_CUDA(ctx, cudaMemcpy(p->d, &p->h, sizeof(PldpConsts), cudaMemcpyHostToDevice));
  p->ready = true;
  return WG_OK;
}

int wg_pldp_solve_batch(wg_ctx *ctx, int mem, int B, const wg_pldp_batch *pb)
{
  if (!ctx || !pb || B < 0) return WG_ERR_INVALID;
  PldpHost *p = pldp_of(ctx);
  if (!p->ready) return wg_fail(ctx, WG_ERR_NOT_READY, "wg_pldp_set_constants not called");
  if (B == 0) return WG_OK;
  if (!pb->D || !pb->m || !pb->DPu || !pb->DPx || !pb->ZMPRef || !pb->XkYk || !pb->X) return WG_ERR_INVALID;
  if (pb->dpu_stride <= 0 || pb->dpx_stride <= 0) return WG_ERR_INVALID;
  wg_device_guard guard(ctx->device);
  const int max_iter = pb->max_iterations > 0 ? pb->max_iterations : 4 * PLDP_KMAX;
  const double tol = 1e-8;   // m_tol, PLDPSolver.cpp:66
  wg_pldp_batch d = *pb;
  // launch over instances [b0, b0 + n) of the device-side batch `d`
  auto launch = [&](int b0, int n) -> int {
    int grid = (n + PLDP_WARPS - 1) / PLDP_WARPS;
    // Measured (16 384 cold-started problems, m ~ 68): constraint matrix staged in shared memory, 8 warps/SM: 4.57 ms; read
    // from L2 (the batch's matrices, 290 MB, stream through once per iteration of their own warp), 32 warps/SM at 64
    // registers: 3.53 ms, 2.85 ms with the work counter (24 warps/SM at 80 registers: 3.10 ms).  The solver is a chain of dependent FP64 operations: resident warps
    // hide more than shared memory saves.  WG_PLDP_STAGE=1 restores the staged variant.
    static const int stage = getenv("WG_PLDP_STAGE") ? atoi(getenv("WG_PLDP_STAGE")) : 0;
    int a_cap = stage ? (int)((pb->dpu_stride + 1) & ~1LL) : 0;
    size_t smem = sizeof(double) * (size_t)a_cap * PLDP_WARPS;
    if (smem > 200 * 1024) { a_cap = 0; smem = 0; }
    WG_SMEM_ATTR(ctx, WG_ATTR_PLDP, pldp_kernel, smem);
    const int per_sm = smem ? (int)((227 * 1024) / (smem + sizeof(PldpWarp) * PLDP_WARPS + 1024)) : 8;
    if (grid > ctx->sm_count * (per_sm > 0 ? per_sm : 1)) grid = ctx->sm_count * (per_sm > 0 ? per_sm : 1);
    if (!p->d_next) WG_CUDA(ctx, cudaMalloc(&p->d_next, sizeof(int)));
    WG_CUDA(ctx, cudaMemsetAsync(p->d_next, 0, sizeof(int), ctx->stream));
    const size_t o = (size_t)b0;
    wg_prof_start(ctx, WG_K_PLDP);
    pldp_kernel<<<grid, PLDP_WARPS * 32, smem, ctx->stream>>>(
        n, p->d, d.D + o * PLDP_U, d.m + o, d.DPu + o * d.dpu_stride, d.dpu_stride, d.DPx + o * d.dpx_stride, d.dpx_stride,
        d.ZMPRef + o * PLDP_U, d.XkYk + o * 6, d.X + o * PLDP_U, d.similar ? d.similar + o * d.similar_stride : nullptr,
        d.similar_stride, d.n_removed ? d.n_removed + o : nullptr, d.starting ? d.starting + o : nullptr,
        d.hot ? d.hot + o : nullptr, pb->hot_start, max_iter, tol, d.info ? d.info + o : nullptr, a_cap, p->d_next);
    wg_prof_stop(ctx);
    WG_LAUNCHED(ctx);
    return WG_OK;
  };
  if (mem == WG_MEM_DEVICE) return launch(0, B);
  if (mem != WG_MEM_HOST) return WG_ERR_INVALID;
  // ---- host buffers: stage, then up to 4 chunks pipelined over upload / kernel / download streams (the upload of the
  // constraint matrices, 17 KB per problem, is twice the kernel time)
  const size_t nb = (size_t)B;
  struct Item { int slot; const void *src; size_t per; const void **dst; };
  const Item in[] = {
      {0, pb->D, sizeof(double) * PLDP_U, (const void **)&d.D},
      {1, pb->m, sizeof(int), (const void **)&d.m},
      {2, pb->DPu, sizeof(double) * (size_t)pb->dpu_stride, (const void **)&d.DPu},
      {3, pb->DPx, sizeof(double) * (size_t)pb->dpx_stride, (const void **)&d.DPx},
      {4, pb->ZMPRef, sizeof(double) * PLDP_U, (const void **)&d.ZMPRef},
      {5, pb->XkYk, sizeof(double) * 6, (const void **)&d.XkYk},
      {6, pb->similar, pb->similar ? sizeof(int) * (size_t)pb->similar_stride : 0, (const void **)&d.similar},
      {7, pb->n_removed, pb->n_removed ? sizeof(int) : 0, (const void **)&d.n_removed},
      {8, pb->starting, pb->starting ? sizeof(int) : 0, (const void **)&d.starting}};
  for (const Item &it : in) {
    if (!it.src) continue;
    int rc = ensure(ctx, p, it.slot, it.per * nb);
    if (rc != WG_OK) return rc;
    *it.dst = p->buf[it.slot];
  }
  // outputs + hot-start state share one allocation: X | info | hot
  const size_t ox = 0, oi = ox + sizeof(double) * PLDP_U * nb, oh = oi + sizeof(wg_pldp_info) * nb;
  int rc = ensure(ctx, p, 9, oh + sizeof(wg_pldp_state) * nb);
  if (rc != WG_OK) return rc;
  char *base = static_cast<char *>(p->buf[9]);
  d.X = reinterpret_cast<double *>(base + ox);
  d.info = pb->info ? reinterpret_cast<wg_pldp_info *>(base + oi) : nullptr;
  d.hot = pb->hot ? reinterpret_cast<wg_pldp_state *>(base + oh) : nullptr;
  if (!p->up) {
    WG_CUDA(ctx, cudaStreamCreateWithFlags(&p->up, cudaStreamNonBlocking));
    WG_CUDA(ctx, cudaStreamCreateWithFlags(&p->down, cudaStreamNonBlocking));
    for (int c = 0; c < 4; ++c) {
      WG_CUDA(ctx, cudaEventCreateWithFlags(&p->ev_up[c], cudaEventDisableTiming));
      WG_CUDA(ctx, cudaEventCreateWithFlags(&p->ev_k[c], cudaEventDisableTiming));
    }
    WG_CUDA(ctx, cudaEventCreateWithFlags(&p->ev0, cudaEventDisableTiming));
  }
  const int nch = std::max(1, std::min(4, B / 2048));
  WG_CUDA(ctx, cudaEventRecord(p->ev0, ctx->stream));
  WG_CUDA(ctx, cudaStreamWaitEvent(p->up, p->ev0, 0));
  for (int c = 0; c < nch; ++c) {
    const size_t b0 = nb * c / nch, b1 = nb * (c + 1) / nch, n = b1 - b0;
    for (const Item &it : in)
      if (it.src)
        WG_CUDA(ctx, cudaMemcpyAsync(static_cast<char *>(p->buf[it.slot]) + b0 * it.per, static_cast<const char *>(it.src) + b0 * it.per,
                                     it.per * n, cudaMemcpyHostToDevice, p->up));
    if (pb->hot) WG_CUDA(ctx, cudaMemcpyAsync(d.hot + b0, pb->hot + b0, sizeof(wg_pldp_state) * n, cudaMemcpyHostToDevice, p->up));
    WG_CUDA(ctx, cudaEventRecord(p->ev_up[c], p->up));
    WG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p->ev_up[c], 0));
    if ((rc = launch((int)b0, (int)n)) != WG_OK) return rc;
    WG_CUDA(ctx, cudaEventRecord(p->ev_k[c], ctx->stream));
    WG_CUDA(ctx, cudaStreamWaitEvent(p->down, p->ev_k[c], 0));
    WG_CUDA(ctx, cudaMemcpyAsync(pb->X + b0 * PLDP_U, d.X + b0 * PLDP_U, sizeof(double) * PLDP_U * n, cudaMemcpyDeviceToHost, p->down));
    if (pb->info) WG_CUDA(ctx, cudaMemcpyAsync(pb->info + b0, d.info + b0, sizeof(wg_pldp_info) * n, cudaMemcpyDeviceToHost, p->down));
    if (pb->hot) WG_CUDA(ctx, cudaMemcpyAsync(pb->hot + b0, d.hot + b0, sizeof(wg_pldp_state) * n, cudaMemcpyDeviceToHost, p->down));
  }
  WG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  WG_CUDA(ctx, cudaStreamSynchronize(p->down));
  return WG_OK;
}

int wg_optcholesky_add_rows_batch(wg_ctx *ctx, int mem, int B, int mode, int nb_max, int card_u, int nb_constraints,
                                  const double *A, long long a_stride, const int *rows, int rows_stride, int k0, int k1,
                                  double *L, long long l_stride)
{
  if (!ctx || B < 0 || nb_max <= 0 || card_u <= 0 || k0 < 0 || k1 < k0 || k1 > nb_max || (mode != 0 && mode != 1))
    return WG_ERR_INVALID;
  if (B == 0 || k1 == k0) return WG_OK;
  if (!A || !rows || !L || rows_stride < k1) return WG_ERR_INVALID;   // UpdateCholeskyMatrix* returns -1 on null A / L
  wg_device_guard guard(ctx->device);
  PldpHost *p = pldp_of(ctx);
  const double *dA = A; const int *drows = rows; double *dL = L;
  const size_t nb = (size_t)B;
  if (mem == WG_MEM_HOST) {
    int rc;
    if ((rc = ensure(ctx, p, 2, sizeof(double) * (size_t)a_stride * nb)) != WG_OK) return rc;
    if ((rc = ensure(ctx, p, 6, sizeof(int) * (size_t)rows_stride * nb)) != WG_OK) return rc;
    if ((rc = ensure(ctx, p, 9, sizeof(double) * (size_t)l_stride * nb)) != WG_OK) return rc;
    WG_CUDA(ctx, cudaMemcpyAsync(p->buf[2], A, sizeof(double) * (size_t)a_stride * nb, cudaMemcpyHostToDevice, ctx->stream));
    WG_CUDA(ctx, cudaMemcpyAsync(p->buf[6], rows, sizeof(int) * (size_t)rows_stride * nb, cudaMemcpyHostToDevice, ctx->stream));
    WG_CUDA(ctx, cudaMemcpyAsync(p->buf[9], L, sizeof(double) * (size_t)l_stride * nb, cudaMemcpyHostToDevice, ctx->stream));
    dA = static_cast<const double *>(p->buf[2]); drows = static_cast<const int *>(p->buf[6]); dL = static_cast<double *>(p->buf[9]);
  } else if (mem != WG_MEM_DEVICE) {
    return WG_ERR_INVALID;
  }
  wg_prof_start(ctx, WG_K_OPTCHOL);
  optchol_kernel<<<grid_for(ctx, B, 4), 128, 0, ctx->stream>>>(B, mode, nb_max, card_u, nb_constraints, dA, a_stride, drows,
                                                               rows_stride, k0, k1, dL, l_stride);
  wg_prof_stop(ctx);
  WG_LAUNCHED(ctx);
  if (mem == WG_MEM_HOST) {
    WG_CUDA(ctx, cudaMemcpyAsync(L, dL, sizeof(double) * (size_t)l_stride * nb, cudaMemcpyDeviceToHost, ctx->stream));
    WG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return WG_OK;
}

int wg_optcholesky_full_batch(wg_ctx *ctx, int mem, int B, int n, const double *A, double *L, double *iL, int inv_size)
{
  if (!ctx || B < 0 || n <= 0 || !L || inv_size < 0 || inv_size > n) return WG_ERR_INVALID;
  if (B == 0) return WG_OK;
  wg_device_guard guard(ctx->device);
  PldpHost *p = pldp_of(ctx);
  const size_t bytes = sizeof(double) * (size_t)n * n * B;
  const double *dA = A; double *dL = L, *diL = iL;
  if (mem == WG_MEM_HOST) {
    int rc;
    if ((rc = ensure(ctx, p, 2, bytes)) != WG_OK) return rc;
    if ((rc = ensure(ctx, p, 9, 2 * bytes)) != WG_OK) return rc;
    if (A) WG_CUDA(ctx, cudaMemcpyAsync(p->buf[2], A, bytes, cudaMemcpyHostToDevice, ctx->stream));
    dL = static_cast<double *>(p->buf[9]); diL = iL ? dL + (size_t)n * n * B : nullptr;
    WG_CUDA(ctx, cudaMemcpyAsync(dL, L, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (iL) WG_CUDA(ctx, cudaMemcpyAsync(diL, iL, bytes, cudaMemcpyHostToDevice, ctx->stream));
    dA = A ? static_cast<const double *>(p->buf[2]) : nullptr;
  } else if (mem != WG_MEM_DEVICE) {
    return WG_ERR_INVALID;
  }
  wg_prof_start(ctx, WG_K_OPTCHOL);
  optchol_full_kernel<<<grid_for(ctx, B, 4), 128, 0, ctx->stream>>>(B, n, dA, dL, diL, inv_size);
  wg_prof_stop(ctx);
  WG_LAUNCHED(ctx);
  if (mem == WG_MEM_HOST) {
    WG_CUDA(ctx, cudaMemcpyAsync(L, dL, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (iL) WG_CUDA(ctx, cudaMemcpyAsync(iL, diL, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    WG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return WG_OK;
}

}  // extern "C"
