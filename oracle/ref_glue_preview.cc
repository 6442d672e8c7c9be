/* oracle/ref_glue_preview.cc - TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" handles onto the reference's own PreviewControl / OptimalControllerSolver object code
 * (compiled by oracle/Makefile from /root/reference/src/PreviewControl/{PreviewControl,
 * OptimalControllerSolver}.cpp over the stand-in MAL header of oracle/ref_shim).  Nothing here restates
 * an algorithm: every function forwards to the reference method named in its comment.
 *
 * LAPACK: OptimalControllerSolver.cpp:44-52 declares dgges_, dlapy2_ and dlamch_ and expects the system
 * LAPACK to provide them; jrl-mal's MAL_INVERSE needs dgetrf_/dgetri_.  This image has no system LAPACK,
 * but the OpenBLAS bundled with the SciPy / OpenCV wheels exports them.  The definitions below
 * are trampolines that forward to the library opened by ref_lapack_open(path) (symbols `dgges_` or
 * `scipy_dgges_`), so libwalkgen_ref.so links without LAPACK and only gain computation needs it.
 */
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <vector>
using std::string;

/* the gains are private members with no getter (PreviewControl.hh:140-160); the glue reads and writes them
 * directly so that a test can (i) fetch the reference's own gains at full precision and (ii) run the
 * reference's recursion on gains of its choosing (ReadPrecomputedFile goes through `float`, :157-176) */
#define private public
#include <PreviewControl/PreviewControl.hh>
#undef private

using namespace PatternGeneratorJRL;

namespace {
void *g_lapack = 0;
void *lapack_sym(const char *name)
{
  if (!g_lapack) return 0;
  void *p = dlsym(g_lapack, name);
  if (!p) {
    std::string s = std::string("scipy_") + name;
    p = dlsym(g_lapack, s.c_str());
  }
  return p;
}
}  // namespace

extern "C" {

int ref_lapack_open(const char *path)
{
  if (g_lapack) return 0;
  g_lapack = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!g_lapack) return -1;
  return (lapack_sym("dgges_") && lapack_sym("dlapy2_") && lapack_sym("dlamch_") && lapack_sym("dgetrf_") &&
          lapack_sym("dgetri_")) ? 0 : -2;
}

typedef long int logical;
typedef logical (*L_fp)(...);

double dlapy2_(double *x, double *y)
{
  typedef double (*fn)(double *, double *);
  return ((fn)lapack_sym("dlapy2_"))(x, y);
}
double dlamch_(char *c)
{
  typedef double (*fn)(char *);
  return ((fn)lapack_sym("dlamch_"))(c);
}
int dgges_(char *jobvsl, char *jobvsr, char *sort, L_fp selctg, int *n, double *a, int *lda, double *b, int *ldb,
           int *sdim, double *alphar, double *alphai, double *beta, double *vsl, int *ldvsl, double *vsr, int *ldvsr,
           double *work, int *lwork, logical *bwork, int *info)
{
  typedef int (*fn)(char *, char *, char *, L_fp, int *, double *, int *, double *, int *, int *, double *, double *,
                    double *, double *, int *, double *, int *, double *, int *, logical *, int *);
  fn f = (fn)lapack_sym("dgges_");
  if (!f) { *info = -999; return 0; }
  return f(jobvsl, jobvsr, sort, selctg, n, a, lda, b, ldb, sdim, alphar, alphai, beta, vsl, ldvsl, vsr, ldvsr, work,
           lwork, bwork, info);
}

}  // extern "C"

/* MAL_INVERSE of the stand-in header: LU inverse through dgetrf_/dgetri_ (row-major storage: the inverse of the
 * transpose is the transpose of the inverse, so the data block can be handed over as is). */
void oracle_mal::lapack_inverse(const oracle_mal::matrix<double> &A, oracle_mal::matrix<double> &invA)
{
  typedef void (*getrf_t)(int *, int *, double *, int *, int *, int *);
  typedef void (*getri_t)(int *, double *, int *, int *, double *, int *, int *);
  getrf_t getrf = (getrf_t)lapack_sym("dgetrf_");
  getri_t getri = (getri_t)lapack_sym("dgetri_");
  int n = (int)A.size1(), info = 0, lwork = 64 * n;
  invA = A;
  if (!getrf || !getri) { invA.fill(0.0 / 0.0); return; }
  std::vector<int> ipiv(n);
  std::vector<double> work(lwork);
  getrf(&n, &n, invA.data(), &n, &ipiv[0], &info);
  if (info == 0) getri(&n, invA.data(), &n, &ipiv[0], &work[0], &lwork, &info);
  if (info != 0) invA.fill(0.0 / 0.0);
}

extern "C" {

/* ---- PreviewControl (PreviewControl.hh:58-140) ---- */
struct RefPreview {
  SimplePluginManager spm;
  PreviewControl *pc;
};

void *ref_preview_new(unsigned mode, int auto_weights)
{
  RefPreview *h = new RefPreview;
  h->pc = new PreviewControl(&h->spm, mode, auto_weights != 0);
  return h;
}
void ref_preview_delete(void *h)
{
  RefPreview *r = static_cast<RefPreview *>(h);
  delete r->pc;
  delete r;
}
/* ReadPrecomputedFile (PreviewControl.cpp:142-196) */
void ref_preview_read_file(void *h, const char *path) { static_cast<RefPreview *>(h)->pc->ReadPrecomputedFile(path); }
/* SetSamplingPeriod / SetPreviewControlTime / SetHeightOfCoM (:103-137) then ComputeOptimalWeights(mode) (:198-322):
 * the reference's own gain computation (needs ref_lapack_open). */
void ref_preview_compute_weights(void *h, double T, double preview_time, double zc, unsigned mode)
{
  PreviewControl *pc = static_cast<RefPreview *>(h)->pc;
  pc->SetSamplingPeriod(T);
  pc->SetPreviewControlTime(preview_time);
  pc->SetHeightOfCoM(zc);
  pc->ComputeOptimalWeights(mode);
}
/* the same through the plugin commands (CallMethod, :512-560) */
void ref_preview_call_method(void *h, const char *method, const char *args)
{
  std::string m(method);
  std::istringstream is(args);
  static_cast<RefPreview *>(h)->pc->CallMethod(m, is);
}
/* private members out: returns m_SizeOfPreviewWindow */
int ref_preview_get_gains(void *h, double *A9, double *B3, double *C3, double *Kx3, double *Ks, double *F, int capF,
                          double *T_Tprev_zc)
{
  PreviewControl *pc = static_cast<RefPreview *>(h)->pc;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) A9[3 * i + j] = pc->m_A(i, j);
    B3[i] = pc->m_B(i, 0);
    C3[i] = pc->m_C(0, i);
    Kx3[i] = pc->m_Kx(0, i);
  }
  *Ks = pc->m_Ks;
  int nl = (int)pc->m_SizeOfPreviewWindow;
  for (int i = 0; i < nl && i < capF && i < (int)pc->m_F.size1(); ++i) F[i] = pc->m_F(i, 0);
  if (T_Tprev_zc) {
    T_Tprev_zc[0] = pc->SamplingPeriod();
    T_Tprev_zc[1] = pc->PreviewControlTime();
    T_Tprev_zc[2] = pc->GetHeightOfCoM();
  }
  return nl;
}
/* private members in: A, B, C are set by the reference itself (ComputeOptimalWeights with a zero preview time
 * fills them and returns early, :202-226); Kx, Ks, F and the window size are written directly. */
void ref_preview_set_gains(void *h, double T, double preview_time, double zc, const double *Kx3, double Ks,
                           const double *F, int NL)
{
  PreviewControl *pc = static_cast<RefPreview *>(h)->pc;
  pc->m_SamplingPeriod = T;
  pc->m_PreviewControlTime = 0.0;
  pc->m_Zc = zc;
  pc->ComputeOptimalWeights(OptimalControllerSolver::MODE_WITHOUT_INITIALPOS);
  pc->m_PreviewControlTime = preview_time;
  for (int i = 0; i < 3; ++i) pc->m_Kx(0, i) = Kx3[i];
  pc->m_Ks = Ks;
  pc->m_SizeOfPreviewWindow = (unsigned)NL;
  MAL_MATRIX_RESIZE(pc->m_F, NL, 1);
  for (int i = 0; i < NL; ++i) pc->m_F(i, 0) = F[i];
}

/* OneIterationOfPreview (:324-374) iterated over one trajectory exactly as the callers do
 * (ZMPPreviewControlWithMultiBodyZMP.cpp:393-403: one call at lindex 0, then pop_front): zmpref_xy [L][2],
 * state8 = {x[3], y[3], sxzmp, syzmp} in/out, com_out [steps][6], zmp_out [steps][2].  Returns the number of steps,
 * or -1 when the reference throws (window longer than the deque). */
long ref_preview_run(void *h, const double *zmpref_xy, long L, double *state8, double *com_out, double *zmp_out,
                     int simulation, int use_lindex)
{
  PreviewControl *pc = static_cast<RefPreview *>(h)->pc;
  std::deque<ZMPPosition> q(L);
  for (long i = 0; i < L; ++i) {
    std::memset(&q[i], 0, sizeof(ZMPPosition));
    q[i].px = zmpref_xy[2 * i];
    q[i].py = zmpref_xy[2 * i + 1];
  }
  MAL_MATRIX_DIM(x, double, 3, 1);
  MAL_MATRIX_DIM(y, double, 3, 1);
  for (int i = 0; i < 3; ++i) { x(i, 0) = state8[i]; y(i, 0) = state8[3 + i]; }
  double sx = state8[6], sy = state8[7], zx = 0, zy = 0;
  const long NL = (long)pc->m_SizeOfPreviewWindow;
  long steps = 0;
  try {
    if (L < NL) {               /* let the reference report it */
      pc->OneIterationOfPreview(x, y, sx, sy, q, 0, zx, zy, simulation != 0);
      return -2;
    }
    for (long k = 0; k + NL <= L; ++k) {
      if (use_lindex) pc->OneIterationOfPreview(x, y, sx, sy, q, (unsigned)k, zx, zy, simulation != 0);
      else {
        pc->OneIterationOfPreview(x, y, sx, sy, q, 0, zx, zy, simulation != 0);
        q.pop_front();
      }
      if (com_out)
        for (int i = 0; i < 3; ++i) { com_out[6 * k + i] = x(i, 0); com_out[6 * k + 3 + i] = y(i, 0); }
      if (zmp_out) { zmp_out[2 * k] = zx; zmp_out[2 * k + 1] = zy; }
      ++steps;
    }
  } catch (...) {
    return -1;
  }
  for (int i = 0; i < 3; ++i) { state8[i] = x(i, 0); state8[3 + i] = y(i, 0); }
  state8[6] = sx; state8[7] = sy;
  return steps;
}

/* The same over walks [b0, b1) of a ragged batch (bench.py --impl reference: the reference's own per-tick call in
 * its own data layout, one walk after the other): offsets [B+1], zmpref_xy [offsets[B]][2], state [B][8],
 * com_out [offsets[B]][6] / zmp_out [offsets[B]][2] (row offsets[b]+k = step k of walk b) or NULL.
 * Walks b0, b0+stride, ... are processed.  Returns the number of preview steps, or -1. */
long ref_preview_run_batch(void *h, int b0, int b1, int stride, const long long *offsets, const double *zmpref_xy,
                           double *state, double *com_out, double *zmp_out, int simulation)
{
  long total = 0;
  for (int b = b0; b < b1; b += stride) {
    const long long o = offsets[b];
    const long L = (long)(offsets[b + 1] - o);
    if (L < (long)static_cast<RefPreview *>(h)->pc->m_SizeOfPreviewWindow) continue;
    const long n = ref_preview_run(h, zmpref_xy + 2 * o, L, state + 8 * (long)b, com_out ? com_out + 6 * o : 0,
                                   zmp_out ? zmp_out + 2 * o : 0, simulation, 0);
    if (n < 0) return -1;
    total += n;
  }
  return total;
}

/* OneIterationOfPreview1D, deque<double> overload (:376-421), iterated with lindex = k. */
long ref_preview_run_1d_deque(void *h, const double *zmpref, long L, double *state4, double *com_out, double *zmp_out,
                              int simulation)
{
  PreviewControl *pc = static_cast<RefPreview *>(h)->pc;
  std::deque<double> q(zmpref, zmpref + L);
  MAL_MATRIX_DIM(x, double, 3, 1);
  for (int i = 0; i < 3; ++i) x(i, 0) = state4[i];
  double sx = state4[3], zx = 0;
  const long NL = (long)pc->m_SizeOfPreviewWindow;
  if (L < NL) return -1;        /* the reference calls exit(0) here (:394-399) */
  long steps = 0;
  for (long k = 0; k + NL <= L; ++k, ++steps) {
    pc->OneIterationOfPreview1D(x, sx, q, (unsigned)k, zx, simulation != 0);
    if (com_out) for (int i = 0; i < 3; ++i) com_out[3 * k + i] = x(i, 0);
    if (zmp_out) zmp_out[k] = zx;
  }
  for (int i = 0; i < 3; ++i) state4[i] = x(i, 0);
  state4[3] = sx;
  return steps;
}

/* OneIterationOfPreview1D, vector<double> overload (:423-484): ONE call at `lindex` on a buffer of L samples,
 * including the wrap-around branch taken when L - lindex < NL (:448-466). */
int ref_preview_step_1d_vector(void *h, const double *zmpref, long L, unsigned lindex, double *x3, double *sxzmp,
                               double *zmpx2, int simulation)
{
  PreviewControl *pc = static_cast<RefPreview *>(h)->pc;
  if (L < (long)pc->m_SizeOfPreviewWindow) return -1;   /* exit(0) in the reference (:439-444) */
  std::vector<double> v(zmpref, zmpref + L);
  MAL_MATRIX_DIM(x, double, 3, 1);
  for (int i = 0; i < 3; ++i) x(i, 0) = x3[i];
  int rc = pc->OneIterationOfPreview1D(x, *sxzmp, v, lindex, *zmpx2, simulation != 0);
  for (int i = 0; i < 3; ++i) x3[i] = x(i, 0);
  return rc;
}

} /* extern "C" */
