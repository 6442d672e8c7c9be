/* oracle/oracle_pldp.cpp - TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Restatement of
 *   OptCholesky::UpdateCholeskyMatrixNormal / Fortran        src/Mathematics/OptCholesky.cpp:123-223
 *   OptCholesky::ComputeNormalCholeskyOnANormal              src/Mathematics/OptCholesky.cpp:225-259
 *   OptCholesky::ComputeInverseCholeskyNormal                src/Mathematics/OptCholesky.cpp:261-302
 *   PLDPSolver::PrecomputeiPuPx / ComputeInitialSolution     src/Mathematics/PLDPSolver.cpp:205-340
 *   PLDPSolver::Forward/BackwardSubstitution                 src/Mathematics/PLDPSolver.cpp:342-400
 *   PLDPSolver::ComputeProjectedDescentDirection             src/Mathematics/PLDPSolver.cpp:404-532
 *   PLDPSolver::ComputeAlpha                                 src/Mathematics/PLDPSolver.cpp:534-653
 *   PLDPSolver::SolveProblem / StoreCurrentZMPSolution       src/Mathematics/PLDPSolver.cpp:654-1032
 * in the reference's order of floating-point operations, WITHOUT the 1.3 ms wall-clock cap (:68-69, :890-900; it makes
 * the reference itself nondeterministic) and with status codes where the reference prints or calls exit(0).
 * Pinned by tests/test_pldp_oracle.py against the reference's own object code (oracle/_ref) on identical inputs
 * (bitwise equal X and identical activation sequences whenever the reference's cap does not bind).
 * Parity unpinned by golden vectors: the reference ships no test or datref for PLDP.
 */
#include <cmath>
#include <cstring>
#include <vector>

namespace {

struct Chol {  /* OptCholesky with caller-owned A / L, MODE_FORTRAN or MODE_NORMAL */
  int nbmax, cardu, mode, nbc = 0;
  const double *A = nullptr;
  double *L = nullptr;
  std::vector<unsigned> act;
  void add(unsigned row)
  {
    act.push_back(row);
    const int i = (int)act.size() - 1;
    const long rs = mode ? 1 : cardu, cs = mode ? (nbc + 1) : 1;
    const double *ri = A + (long)act[i] * rs;
    for (int lj = 0; lj < (int)act.size(); ++lj) {
      const double *rj = A + (long)act[lj] * rs;
      double Mij = 0.0;
      for (int lk = 0; lk < cardu; ++lk) Mij += ri[lk * cs] * rj[lk * cs];
      double r = Mij;
      for (int lk = 0; lk < lj; ++lk) r = r - L[(long)i * nbmax + lk] * L[(long)lj * nbmax + lk];
      if (lj != (int)act.size() - 1) L[(long)i * nbmax + lj] = r / L[(long)lj * nbmax + lj];
      else L[(long)i * nbmax + lj] = sqrt(r);
    }
  }
};

}  // namespace

extern "C" {

/* AddActiveConstraint for rows[k0..k1): L row-major, leading dimension nbmax. */
void oracle_optcholesky_add_rows(int mode, int nbmax, int cardu, int nbc, const double *A, const int *rows, int k0,
                                 int k1, double *L)
{
  Chol c{nbmax, cardu, mode};
  c.nbc = nbc; c.A = A; c.L = L;
  for (int i = 0; i < k0; ++i) c.act.push_back((unsigned)rows[i]);
  for (int i = k0; i < k1; ++i) c.add((unsigned)rows[i]);
}

void oracle_optcholesky_full(int n, const double *A, double *L)
{
  for (int li = 0; li < n; ++li)
    for (int lj = 0; lj <= li; ++lj) {
      double r = A[(long)li * n + lj];
      for (int lk = 0; lk < lj; ++lk) r = r - L[(long)li * n + lk] * L[(long)lj * n + lk];
      if (lj != li) L[(long)li * n + lj] = r / L[(long)lj * n + lj];
      else L[(long)li * n + lj] = sqrt(r);
    }
}

void oracle_optcholesky_inverse(int n, int size, const double *L, double *iL)
{
  for (int lj = size - 1; lj >= 0; --lj) {
    double inv = 1 / L[(long)lj * n + lj];
    iL[(long)lj * n + lj] = inv;
    for (int li = lj + 1; li < size; ++li) {
      double r = 0.0;
      for (int lk = lj + 1; lk < size; ++lk) r = r + iL[(long)li * n + lk] * L[(long)lk * n + lj];
      iL[(long)li * n + lj] = -inv * r;
    }
  }
}

struct OraclePldpState {
  double prev_zmp[32];
  int prev_active[32];
  int n_prev;
  int pad_;
};

/* One SolveProblem.  info[0..3] = rc, status, iterations, n_active; active[32].
 * similar: SimilarConstraints [m] or NULL.  ComputeAlpha's reuse (PLDPSolver.cpp:570-590) is restated as written:
 * m_ConstraintsValueComputed[li] is reset when row li is visited and set once its product is formed, so a flag that
 * points BACKWARD (li + similar[li] < li: the only kind FindSimilarConstraints produces, -2 / -3) reuses -tmp1 of that row
 * iff the row is not active; the second reuse (:600-610) tests m_ConstraintsValueComputed[lindex + m], which nothing ever
 * sets, and is dead code.  A flag pointing forward would read the flag array as the PREVIOUS ComputeAlpha call left it
 * (uninitialised memory on the first one): status 6 here. */
int oracle_pldp_solve_sim(int N, const double *iPu, const double *Px, const double *Pu, const double *D, int m,
                          const double *A, const double *b, const double *ZMPRef, const double *XkYk, double *X,
                          int n_removed, int starting, OraclePldpState *hot, int hot_start, int max_iter, int *info,
                          int *active_out, const int *similar)
{
  const int U = 2 * N, ld = m + 1;
  const double tol = 1e-8;
  std::vector<double> iPuPx(U * 6, 0.0), Vk(U), c(U), d(U), L(32 * 32, 0.0), v1(32), v2(32, 0.0), y(32), tmp1(m + 1), tmp2(m + 1);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < N; ++k) {
        double tmp = iPu[k * N + i] * Px[k * 3 + j];
        iPuPx[i * 6 + j] += tmp;
        iPuPx[(i + N) * 6 + j + 3] += tmp;
      }
  const bool hs = hot && hot_start;
  for (int i = 0; i < N; ++i) {
    Vk[i] = 0.0; Vk[i + N] = 0.0;
    for (int j = 0; j < 3; ++j) Vk[i] -= iPuPx[i * 6 + j] * XkYk[j];
    for (int j = 3; j < 6; ++j) Vk[i + N] -= iPuPx[(i + N) * 6 + j] * XkYk[j];
    if (hs && !starting) {
      for (int j = 0; j < N - 1; ++j) Vk[i] += iPu[j * N + i] * hot->prev_zmp[j + 1];
      Vk[i] += iPu[(N - 1) * N + i] * ZMPRef[N - 1];
      for (int j = 0; j < N - 1; ++j) Vk[i + N] += iPu[j * N + i] * hot->prev_zmp[j + N + 1];
      Vk[i + N] += iPu[(N - 1) * N + i] * ZMPRef[N - 1 + N];
    } else {
      for (int j = 0; j < N; ++j) Vk[i] += iPu[j * N + i] * ZMPRef[j];
      for (int j = 0; j < N; ++j) Vk[i + N] += iPu[j * N + i] * ZMPRef[j + N];
    }
  }
  Chol ch{32, U, 1};
  ch.nbc = m; ch.A = A; ch.L = L.data();
  std::vector<unsigned> act;
  int status = 0;
  if (hs)
    for (int i = 0; i < hot->n_prev && (int)act.size() < 32; ++i) {
      int idx = hot->prev_active[i] - n_removed;
      if (idx >= 0 && idx < m) { act.push_back(idx); ch.add(idx); }
    }
  int it = 0;
  bool cont = true;
  size_t kproj = 0;
  std::vector<char> computed(2 * (size_t)m + 2, 0);
  while (cont) {
    for (int i = 0; i < U; ++i) c[i] = -D[i] - Vk[i];
    const size_t k = act.size();
    for (size_t li = 0; li < k; ++li) {
      v1[li] = 0.0;
      for (int lj = 0; lj < U; ++lj) v1[li] += A[act[li] + (long)lj * ld] * c[lj];
    }
    for (size_t i = 0; i < k; ++i) {
      y[i] = v1[i];
      for (size_t kk = 0; kk < i; ++kk) y[i] += -L[i * 32 + kk] * y[kk];
      if (L[i * 32 + i] != 0.0) y[i] /= L[i * 32 + i];
    }
    for (int i = (int)k - 1; i >= 0; --i) {
      v2[i] = y[i];
      for (int kk = i + 1; kk < (int)k; ++kk) v2[i] -= L[kk * 32 + i] * v2[kk];
      v2[i] = v2[i] / L[i * 32 + i];
    }
    kproj = k;
    for (int li = 0; li < U; ++li) {
      d[li] = c[li];
      for (size_t lj = 0; lj < k; ++lj) d[li] -= A[act[lj] + (long)li * ld] * v2[lj];
    }
    double Alpha = 10000000.0;
    bool toadd = false;
    unsigned which = 0;
    for (int li = 0; li < m; ++li) {
      bool found = false;
      computed[li] = 0; computed[li + m] = 0;
      for (size_t q = 0; q < k; ++q) if ((int)act[q] == li) { found = true; break; }
      if (found) continue;
      tmp1[li] = 0.0;
      bool tbc = true;
      if (similar && similar[li] != 0) {
        const int lindex = li + similar[li];
        if (lindex < 0 || lindex >= li) { status = 6; cont = false; break; }
        if (computed[lindex]) { tmp1[li] = -tmp1[lindex]; tbc = false; }
      }
      if (tbc)
        for (int lj = 0; lj < U; ++lj) tmp1[li] += A[li + (long)lj * ld] * d[lj];
      computed[li] = 1;
      if (tmp1[li] < 0.0) {
        tmp2[li] = -b[li];
        tbc = true;
        if (similar && similar[li] != 0 && computed[li + similar[li] + m]) {   /* never true, :600-610 */
          tmp2[li] += -tmp2[li + similar[li]] - b[li + similar[li]];
          tbc = false;
        }
        if (tbc)
          for (int lj = 0; lj < U; ++lj) tmp2[li] -= A[li + (long)lj * ld] * Vk[lj];
        if (tmp2[li] > tol) status = status > 1 ? status : 1;
        else if (tmp2[li] > 0.0) tmp2[li] = -tol;
        double la = tmp2[li] / tmp1[li];
        if (Alpha > la) { Alpha = la; if (Alpha < 1) { toadd = true; which = li; } }
      }
    }
    if (status == 6) break;
    double alpha = Alpha;
    if (alpha >= 1.0) { alpha = 1.0; cont = false; }
    if (alpha < 0.0) { status = 2; cont = false; }
    if (status != 2) for (int i = 0; i < U; ++i) Vk[i] = Vk[i] + alpha * d[i];
    if (cont) {
      if (act.size() >= 32 || !toadd) { status = 3; cont = false; }
      else { act.push_back(which); ch.add(which); }
    }
    ++it;
    if (it >= max_iter && cont) { cont = false; if (!status) status = 4; }
  }
  for (int i = 0; i < U; ++i) X[i] = Vk[i];
  if (hs) {
    hot->n_prev = 0;
    for (size_t i = 0; i < kproj; ++i) if (v2[i] < 0.0) hot->prev_active[hot->n_prev++] = (int)act[i];
    for (int i = 0; i < N; ++i) {
      hot->prev_zmp[i] = 0.0; hot->prev_zmp[i + N] = 0.0;
      for (int j = 0; j < N; ++j) { hot->prev_zmp[i] += Pu[j * N + i] * Vk[j]; hot->prev_zmp[i + N] += Pu[j * N + i] * Vk[j + N]; }
      for (int j = 0; j < 3; ++j) { hot->prev_zmp[i] += Px[i * 3 + j] * XkYk[j]; hot->prev_zmp[i + N] += Px[i * 3 + j] * XkYk[j + 3]; }
    }
  }
  int rc = 0;
  if (std::isnan(X[0]) || std::isnan(X[N]) || std::isinf(X[0]) || std::isinf(X[N])) rc = -1;
  if (info) { info[0] = rc; info[1] = status; info[2] = it; info[3] = (int)act.size(); }
  if (active_out) for (int i = 0; i < 32; ++i) active_out[i] = i < (int)act.size() ? (int)act[i] : -1;
  return rc;
}

int oracle_pldp_solve(int N, const double *iPu, const double *Px, const double *Pu, const double *D, int m,
                      const double *A, const double *b, const double *ZMPRef, const double *XkYk, double *X,
                      int n_removed, int starting, OraclePldpState *hot, int hot_start, int max_iter, int *info,
                      int *active_out)
{
  return oracle_pldp_solve_sim(N, iPu, Px, Pu, D, m, A, b, ZMPRef, XkYk, X, n_removed, starting, hot, hot_start, max_iter,
                               info, active_out, nullptr);
}

/* B cold-start problems back to back (CPU-baseline loop of bench.py: no Python in the timed region). */
long oracle_pldp_solve_batch(int N, const double *iPu, const double *Px, const double *Pu, int B, const double *D,
                             const int *m, const double *DPu, long dpu_stride, const double *DPx, long dpx_stride,
                             const double *ZMPRef, const double *XkYk, double *X, int *iterations)
{
  long fails = 0;
  int info[4];
  for (int b = 0; b < B; ++b) {
    oracle_pldp_solve(N, iPu, Px, Pu, D + (long)b * 2 * N, m[b], DPu + (long)b * dpu_stride, DPx + (long)b * dpx_stride,
                      ZMPRef + (long)b * 2 * N, XkYk + (long)b * 6, X + (long)b * 2 * N, 0, 1, nullptr, 0, 128, info,
                      nullptr);
    if (iterations) iterations[b] = info[2];
    fails += (info[0] != 0 || info[1] != 0);
  }
  return fails;
}

} /* extern "C" */
