/* walkgen_b200.h - C ABI of the B200-native batched ZMP pattern-generation backend.
 *
 * This is the drop-in boundary for the jrl-walkgen hot path.  The reference
 * has no C ABI of its own apart from ql0001_ (src/Mathematics/qld.hh:27-31); every entry point below
 * names the reference interface it replaces.  Conventions:
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - every function returns 0 (WG_OK) or a negative wg_status; none throws or calls exit()
 *     (the reference LTHROWs / exit(0)s on these paths: PreviewControl.cpp:341-344, :394-399,
 *     PLDPSolver.cpp:822-828);
 *   - `mem` says where the *bulk* buffers of a call live: WG_MEM_HOST (the library stages them
 *     through the context's stream: H2D, kernels, D2H) or WG_MEM_DEVICE (used in place, no copies).
 *     Small metadata arrays documented as "host" are always host pointers;
 *   - calls are asynchronous on the context's CUDA stream when mem == WG_MEM_DEVICE; call wg_sync()
 *     before reading results.  With WG_MEM_HOST the call returns after the D2H copy completed;
 *   - all arithmetic is FP64 (the reference is double throughout);
 *   - there is NO CPU fallback: without a CUDA device wg_ctx_create fails with WG_ERR_NO_DEVICE.
 */
#ifndef WALKGEN_B200_H
#define WALKGEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum wg_status {
  WG_OK = 0,
  WG_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product never falls back to the CPU */
  WG_ERR_CUDA = -2,        /* a CUDA runtime call failed; see wg_last_error()                     */
  WG_ERR_INVALID = -3,     /* bad argument                                                        */
  WG_ERR_ALLOC = -4,
  WG_ERR_NOT_READY = -5,   /* e.g. preview gains not set                                          */
  WG_ERR_WINDOW = -6       /* fewer ZMP samples than the preview window (PreviewControl.cpp:341)  */
} wg_status;

enum { WG_MEM_HOST = 0, WG_MEM_DEVICE = 1 };

typedef struct wg_ctx wg_ctx;

/* ------------------------------------------------------------------------------------------------
 * Context, memory, synchronisation
 * ---------------------------------------------------------------------------------------------- */
int wg_version(void);                               /* (major<<16)|(minor<<8)|patch                */
int wg_device_count(void);                          /* 0 when no usable CUDA device                */
int wg_ctx_create(int device, wg_ctx **out);        /* one context per GPU / per process rank      */
int wg_ctx_destroy(wg_ctx *ctx);
int wg_sync(wg_ctx *ctx);                           /* cudaStreamSynchronize on the context stream */
const char *wg_last_error(wg_ctx *ctx);             /* message of the last failure (never NULL)    */
void *wg_ctx_stream(wg_ctx *ctx);                   /* the cudaStream_t, for callers that time it  */
int wg_malloc_device(wg_ctx *ctx, size_t bytes, void **out);
int wg_free_device(wg_ctx *ctx, void *p);
int wg_malloc_pinned(wg_ctx *ctx, size_t bytes, void **out);
int wg_free_pinned(wg_ctx *ctx, void *p);
int wg_memcpy_h2d(wg_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);  /* async */
int wg_memcpy_d2h(wg_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);  /* async */
int wg_memset_device(wg_ctx *ctx, void *dst_dev, int value, size_t bytes);          /* async */
/* CUDA-event stopwatch on the context stream (for bench.py; torch events cannot see this stream). */
int wg_timer_start(wg_ctx *ctx);
int wg_timer_stop_ms(wg_ctx *ctx, float *ms);       /* records, synchronises, returns elapsed ms   */
/* Number of kernels this library launched on this context since creation / last reset. */
long long wg_launch_count(wg_ctx *ctx);
void wg_launch_count_reset(wg_ctx *ctx);
/* FP64 FMA peak micro-benchmark (register-resident DFMA chains on every SM): TFLOP/s. */
int wg_measure_fp64_peak(wg_ctx *ctx, double *tflops);

/* ------------------------------------------------------------------------------------------------
 * Kajita2003 preview control
 *   replaces PreviewControl::ComputeOptimalWeights        (src/PreviewControl/PreviewControl.cpp:198-322)
 *            OptimalControllerSolver::ComputeWeights       (src/PreviewControl/OptimalControllerSolver.cpp:200-352)
 *            PreviewControl::OneIterationOfPreview         (src/PreviewControl/PreviewControl.cpp:324-374)
 *            PreviewControl::OneIterationOfPreview1D (x2)  (src/PreviewControl/PreviewControl.cpp:376-484)
 * ---------------------------------------------------------------------------------------------- */
#define WG_PREVIEW_MAX_NL 2048

enum { WG_PREVIEW_MODE_WITH_INITIALPOS = 0,      /* OptimalControllerSolver.hh:144 */
       WG_PREVIEW_MODE_WITHOUT_INITIALPOS = 1 }; /* OptimalControllerSolver.hh:143 */

typedef struct wg_preview_gains_t {
  double A[9];   /* cart-table state matrix, row-major   (PreviewControl.cpp:204-206) */
  double B[3];   /*                                      (PreviewControl.cpp:208-210) */
  double C[3];   /* zmp = C x, C = [1, 0, -zc/9.81]      (PreviewControl.cpp:212-214) */
  double Kx[3];  /* state feedback                       (PreviewControl.cpp:279-280) */
  double Ks;     /* integral gain                        (PreviewControl.cpp:278)     */
  double T, preview_time, zc;
  int mode;
  int NL;        /* preview window = (int)(preview_time/T) (PreviewControl.cpp:317-319) */
  double F[WG_PREVIEW_MAX_NL];
} wg_preview_gains_t;

/* Host-side Riccati solve (runs once per parameter change, as in the reference).  The reference
 * calls LAPACK dgges_ on the symplectic pencil (Laub); here the same stabilising solution is
 * obtained with the structure-preserving doubling algorithm. */
int wg_preview_gains(double T, double preview_time, double zc, int mode, wg_preview_gains_t *out);

/* Upload gains to the context (constant memory image used by the kernels). */
int wg_preview_set_gains(wg_ctx *ctx, const wg_preview_gains_t *gains);

/* A plan describes a ragged batch of B trajectories: trajectory b owns ZMP-reference samples
 * [offsets[b], offsets[b+1]) of a packed array of interleaved (px,py) pairs.  A trajectory of L
 * samples yields L-NL+1 preview steps (step k consumes the window [k, k+NL), exactly what
 * OneIterationOfPreview reads from the FIFO at lindex=k).  offsets is a HOST array of B+1 entries. */
typedef struct wg_preview_plan wg_preview_plan;
int wg_preview_plan_create(wg_ctx *ctx, int B, const int64_t *offsets_host, wg_preview_plan **out);
int wg_preview_plan_destroy(wg_preview_plan *plan);
int64_t wg_preview_plan_total_steps(const wg_preview_plan *plan);   /* sum of (L_b-NL+1)^+ */
int64_t wg_preview_plan_total_samples(const wg_preview_plan *plan); /* offsets[B]           */

/* Run every preview step of every trajectory of the plan.
 *   zmpref_xy : [total_samples][2]  in      ZMP reference (px,py)
 *   state     : [B][8]              in/out  {x, dx, ddx, y, dy, ddy, sxzmp, syzmp}: the arguments
 *                                           x, y, sxzmp, syzmp of OneIterationOfPreview
 *   com_out   : [total_samples][6]  out     row offsets[b]+k = (x,dx,ddx,y,dy,ddy) after step k
 *   zmp_out   : [total_samples][2]  out     row offsets[b]+k = (zmpx2, zmpy2) of step k
 *                                           (rows past a trajectory's last step: left untouched with
 *                                           WG_MEM_DEVICE, zero with WG_MEM_HOST)
 *   simulation: the `Simulation` flag (accumulate sxzmp += zmpref - zmp, PreviewControl.cpp:363-367)
 * com_out / zmp_out may be NULL. */
int wg_preview_run_batch(wg_ctx *ctx, wg_preview_plan *plan, int mem, const double *zmpref_xy,
                         double *state, double *com_out, double *zmp_out, int simulation);

/* Single-call form of OneIterationOfPreview for the class wrappers (batch of one, one step):
 * host pointers; x[3], y[3], sxzmp, syzmp in/out; window = NL pairs starting at ZMPPositions[lindex]. */
int wg_preview_one_iteration(wg_ctx *ctx, double *x, double *y, double *sxzmp, double *syzmp,
                             const double *window_xy, int n_available, double *zmpx2, double *zmpy2,
                             int simulation);

#ifdef __cplusplus
}
#endif
#endif /* WALKGEN_B200_H */
