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
/* Per-kernel CUDA-event profiler on the context stream.  Between wg_prof_begin(capacity = max number of
 * kernel launches to record) and wg_prof_end every kernel this library launches is bracketed by an event
 * pair; wg_prof_get returns the launch count and the summed duration of one kernel id:
 *   0 preview FIR, 1 preview recursion, 2 Herdt QP solve, 3 Herdt MPC step, 4 PLDP solve, 5 OptCholesky,
 *   6 preview fused. */
int wg_prof_begin(wg_ctx *ctx, int capacity);
int wg_prof_end(wg_ctx *ctx);                        /* synchronises the stream and accumulates      */
int wg_prof_get(wg_ctx *ctx, int kernel_id, long long *launches, double *total_ms);
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

/* ------------------------------------------------------------------------------------------------
 * Herdt2010 velocity-referenced QP (N = 16 previewed samples of T = 0.1 s)
 *   replaces GeneratorVelRef::build_invariant_part / update_problem / build_constraints
 *                (src/ZMPRefTrajectoryGeneration/generator-vel-ref.cpp:588-674, :285-474, :555-584)
 *            RelativeFeetInequalities::set_vertices / compute_linear_system
 *                (src/Mathematics/relative-feet-inequalities.cpp:186-234, :265-319)
 *            QPProblem::add_term_to / solve          (src/ZMPRefTrajectoryGeneration/qp-problem.cpp:411-547, :246-407)
 *            ql0001_ / ql0002_                       (src/Mathematics/qld.cpp:378-2090)
 *            LinearizedInvertedPendulum2D::Interpolation / OneIteration
 *                (src/PreviewControl/LinearizedInvertedPendulum2D.cpp:157-264)
 * ---------------------------------------------------------------------------------------------- */
#define WG_HERDT_N 16            /* QP_N_, ZMPVelocityReferencedQP.cpp:65                        */
#define WG_HERDT_MAX_STEPS 2     /* previewed steps ns in {0,1,2} for StepPeriod = 8 samples     */
#define WG_HERDT_MAX_VARS (2 * WG_HERDT_N + 2 * WG_HERDT_MAX_STEPS)            /* 36 */
#define WG_HERDT_MAX_ROWS (1 + 4 * WG_HERDT_N + 5 * WG_HERDT_MAX_STEPS)        /* 75: dummy row + CoP + feet */

enum { WG_LEFT = 0, WG_RIGHT = 1 };  /* foot_type_e, privatepgtypes.hh:47-50 */
enum { WG_SS = 0, WG_DS = 1 };       /* PhaseType,   privatepgtypes.hh:65-68 */

/* Generator + robot constants (one set per context). */
typedef struct wg_herdt_params {
  double T;                 /* QP sampling period, 0.1          (ZMPVelocityReferencedQP.cpp:63)   */
  double com_height;        /* QP model height, 0.814           (ZMPVelocityReferencedQP.cpp:103)  */
  double w_jerk;            /* JERK_MIN weight 1e-5             (ZMPVelocityReferencedQP.cpp:118)  */
  double w_vel;             /* INSTANT_VELOCITY weight 1.0      (ZMPVelocityReferencedQP.cpp:116)  */
  double w_cop;             /* COP_CENTERING weight 1e-6        (ZMPVelocityReferencedQP.cpp:117)  */
  double cop_half_x;        /* 0.5*sole_length - SecurityMarginX (FootHalfSize.cpp:62-68)          */
  double cop_half_y;        /* 0.5*sole_width  - SecurityMarginY                                   */
  double ds_feet_distance;  /* DSFeetDistance_ 0.2              (relative-feet-inequalities.cpp:45) */
  double foot_hull_x[5];    /* LeftFPosEdgesX_                  (relative-feet-inequalities.cpp:51) */
  double foot_hull_y[5];    /* LeftFPosEdgesY_ (right = -left)  (relative-feet-inequalities.cpp:52) */
  double lipm_T;            /* control period for the 5 ms interpolation, 0.005                    */
} wg_herdt_params;

/* Defaults of the reference for a sole of (length, width) metres and margins 0.04 m. */
void wg_herdt_default_params(double sole_length, double sole_width, wg_herdt_params *out);

/* Everything GeneratorVelRef reads when it assembles one QP (one instance):
 * IntermedData_->State().CoM, Ref.Global.{X,Y}_vec and Solution_.SupportStates_deq[0..N]. */
typedef struct wg_herdt_qp_input {
  double com_x[3], com_y[3];                       /* (c, dc, ddc) per axis                          */
  double ref_x[WG_HERDT_N], ref_y[WG_HERDT_N];     /* velocity reference in the global frame         */
  double sup_x[WG_HERDT_N + 1];                    /* support_state_t::X   [0] = current, [i] = previewed */
  double sup_y[WG_HERDT_N + 1];                    /* support_state_t::Y                             */
  double sup_yaw[WG_HERDT_N + 1];                  /* support_state_t::Yaw                           */
  int8_t sup_foot[WG_HERDT_N + 1];                 /* WG_LEFT / WG_RIGHT                             */
  int8_t sup_phase[WG_HERDT_N + 1];                /* WG_SS / WG_DS                                  */
  int8_t sup_step[WG_HERDT_N + 1];                 /* support_state_t::StepNumber                    */
  int8_t sup_changed[WG_HERDT_N + 1];              /* support_state_t::StateChanged                  */
  int8_t pad_[4];
} wg_herdt_qp_input;                               /* 784 bytes */

/* solution_t as filled by QPProblem::solve (qp-problem.cpp:281-293) plus the first 0.1 s of CoM/ZMP. */
typedef struct wg_herdt_qp_output {
  double x[WG_HERDT_MAX_VARS];      /* Solution_vec: jerk_x[16], jerk_y[16], foot_x[ns], foot_y[ns]       */
  double lagr[WG_HERDT_MAX_ROWS + 1]; /* ConstrLagr_vec; row 0 is the reference's all-zero dummy row      */
  double com_next_x[3], com_next_y[3]; /* CoM_ after OneIteration(x[0], x[N]) (LIPM2D.cpp:230-264)        */
  int32_t n_vars;                   /* 2N + 2ns                                                           */
  int32_t n_rows;                   /* m_ = 1 + 4N + 5ns (incl. dummy row)                                */
  int32_t fail;                     /* 0 ok (QLD ifail convention: >0 failure)                            */
  int32_t iterations;               /* active-set changes (adds + drops)                                  */
} wg_herdt_qp_output;

int wg_herdt_set_params(wg_ctx *ctx, const wg_herdt_params *params);

/* Build and solve B independent QPs.  in/out are arrays of B structs (host or device per `mem`). */
int wg_herdt_qp_solve_batch(wg_ctx *ctx, int mem, int B, const wg_herdt_qp_input *in,
                            wg_herdt_qp_output *out);

#ifdef __cplusplus
}
#endif
#endif /* WALKGEN_B200_H */
