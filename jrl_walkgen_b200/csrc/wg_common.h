// wg_common.h - internal definitions shared by the CUDA translation units of libwalkgen_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/walkgen_b200.h"

#define WG_VERSION ((0 << 16) | (1 << 8) | 0)

struct wg_preview_consts;  // preview.cu

struct wg_ctx {
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  long long launches = 0;
  char err[512] = {0};
  // preview control
  bool preview_ready = false;
  wg_preview_gains_t preview_gains;
  double *d_previewF = nullptr;  // device copy of F padded with zeros
  // Herdt constants (herdt_qp.cu)
  void *herdt = nullptr;
  // PLDP constants (pldp.cu)
  void *pldp = nullptr;
};

inline int wg_fail(wg_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess)
{
  if (ctx) {
    if (e != cudaSuccess)
      snprintf(ctx->err, sizeof ctx->err, "%s: %s", what, cudaGetErrorString(e));
    else
      snprintf(ctx->err, sizeof ctx->err, "%s", what);
  }
  return code;
}

#define WG_CUDA(ctx, call)                                                   \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) return wg_fail((ctx), WG_ERR_CUDA, #call, e__);  \
  } while (0)

#define WG_LAUNCHED(ctx)                                                     \
  do {                                                                       \
    (ctx)->launches++;                                                       \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) return wg_fail((ctx), WG_ERR_CUDA, "kernel launch", e__); \
  } while (0)

struct wg_device_guard {
  int prev = -1;
  explicit wg_device_guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~wg_device_guard() { if (prev >= 0) cudaSetDevice(prev); }
};
