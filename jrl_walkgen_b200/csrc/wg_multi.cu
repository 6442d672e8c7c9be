// wg_multi.cu - library-level multi-GPU: one process drives several B200s of a box (SURVEY 8e, 8b "wg_ctx_create(device_mask)").
//
// Instances of this path are independent, so the data path has no collective: a wg_multi owns one wg_ctx per selected device,
// a sharded run deals instance i to device i mod G, every device runs the identical kernel sequence on its share from its own
// host thread, and ONE collective closes the run: the per-device statistics (solves, failures, active-set changes, instances
// still on line; and the CUDA-event time of the slowest device) are all-reduced over NCCL across the devices.  NCCL is bound at
// run time (dlopen of libnccl.so.2: the library has no link-time NCCL dependency and stays loadable on a box without it, where
// the statistics are summed on the host instead and the result says so).
#include "wg_common.h"
#include <dlfcn.h>
#include <algorithm>
#include <thread>
#include <vector>

namespace {

// the few NCCL entry points used, with the ABI of nccl.h 2.x
typedef void *ncclComm_t;
typedef int ncclResult_t;
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2 };
struct Nccl {
  void *lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  bool ok() const { return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd; }
};

bool nccl_load(Nccl &n)
{
  const char *env = getenv("WG_NCCL_LIB");
  const char *names[3] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm || !nm[0]) continue;
    n.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    if (n.lib) break;
  }
  if (!n.lib) return false;
  n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(dlsym(n.lib, "ncclCommInitAll"));
  n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(n.lib, "ncclCommDestroy"));
  n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(n.lib, "ncclAllReduce"));
  n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(dlsym(n.lib, "ncclGroupStart"));
  n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(dlsym(n.lib, "ncclGroupEnd"));
  n.GetVersion = reinterpret_cast<decltype(n.GetVersion)>(dlsym(n.lib, "ncclGetVersion"));
  return n.ok();
}

}  // namespace

struct wg_multi {
  std::vector<wg_ctx *> ctx;
  std::vector<int> dev;
  std::vector<double *> d_stats;     // 8 doubles per device: [0..3] sums, [4] time for the MAX reduction
  Nccl nccl;
  std::vector<ncclComm_t> comm;
  bool use_nccl = false;
  int nccl_version = 0;
  char err[256] = {0};
};

extern "C" {

int wg_multi_create(unsigned long long device_mask, wg_multi **out)
{
  if (!out) return WG_ERR_INVALID;
  *out = nullptr;
  const int ndev = wg_device_count();
  if (ndev <= 0) return WG_ERR_NO_DEVICE;
  wg_multi *m = new (std::nothrow) wg_multi();
  if (!m) return WG_ERR_ALLOC;
  for (int d = 0; d < ndev && d < 64; ++d)
    if (device_mask == 0 || ((device_mask >> d) & 1ull)) m->dev.push_back(d);
  if (m->dev.empty()) { delete m; return WG_ERR_INVALID; }
  for (int d : m->dev) {
    wg_ctx *c = nullptr;
    const int rc = wg_ctx_create(d, &c);
    if (rc != WG_OK) { for (wg_ctx *x : m->ctx) wg_ctx_destroy(x); delete m; return rc; }
    m->ctx.push_back(c);
  }
  for (size_t k = 0; k < m->ctx.size(); ++k) {
    double *p = nullptr;
    if (wg_malloc_device(m->ctx[k], sizeof(double) * 8, reinterpret_cast<void **>(&p)) != WG_OK) p = nullptr;
    m->d_stats.push_back(p);
  }
  // NCCL communicators over the selected devices (single process: ncclCommInitAll)
  if (nccl_load(m->nccl)) {
    m->comm.assign(m->dev.size(), nullptr);
    if (m->nccl.CommInitAll(m->comm.data(), (int)m->dev.size(), m->dev.data()) == 0) {
      m->use_nccl = true;
      if (m->nccl.GetVersion) m->nccl.GetVersion(&m->nccl_version);
    } else {
      m->comm.clear();
      snprintf(m->err, sizeof m->err, "ncclCommInitAll failed: statistics are reduced on the host");
    }
  } else {
    snprintf(m->err, sizeof m->err, "libnccl.so.2 not found: statistics are reduced on the host");
  }
  *out = m;
  return WG_OK;
}

int wg_multi_destroy(wg_multi *m)
{
  if (!m) return WG_OK;
  if (m->use_nccl) for (ncclComm_t c : m->comm) if (c) m->nccl.CommDestroy(c);
  for (size_t k = 0; k < m->ctx.size(); ++k) {
    if (m->d_stats[k]) wg_free_device(m->ctx[k], m->d_stats[k]);
    wg_ctx_destroy(m->ctx[k]);
  }
  delete m;
  return WG_OK;
}

int wg_multi_size(const wg_multi *m) { return m ? (int)m->ctx.size() : 0; }
wg_ctx *wg_multi_ctx(wg_multi *m, int k) { return (m && k >= 0 && k < (int)m->ctx.size()) ? m->ctx[k] : nullptr; }
int wg_multi_nccl_version(const wg_multi *m) { return (m && m->use_nccl) ? m->nccl_version : 0; }
const char *wg_multi_last_error(const wg_multi *m) { return m ? m->err : ""; }

int wg_multi_herdt_set_params(wg_multi *m, const wg_herdt_params *hp, const wg_herdt_mpc_params *mp)
{
  if (!m || !hp || !mp) return WG_ERR_INVALID;
  for (wg_ctx *c : m->ctx) {
    int rc = wg_herdt_set_params(c, hp);
    if (rc == WG_OK) rc = wg_herdt_mpc_set_params(c, mp);
    if (rc != WG_OK) { snprintf(m->err, sizeof m->err, "device %d: %s", c->device, wg_last_error(c)); return rc; }
  }
  return WG_OK;
}

int wg_multi_herdt_mpc_sweep(wg_multi *m, long long instances, int periods, int chunk, const double *vel_ref, const double *init9,
                             wg_multi_stats *out)
{
  if (!m || instances < 0 || periods < 0 || !vel_ref || !init9 || !out) return WG_ERR_INVALID;
  const int G = (int)m->ctx.size();
  if (chunk <= 0) chunk = 10;
  std::memset(out, 0, sizeof *out);
  out->devices = G;
  std::vector<int> rcs(G, WG_OK);
  std::vector<float> ms(G, 0.f);
  std::vector<long long> launches(G, 0);
  std::vector<void *> d_states(G, nullptr);
  std::vector<long long> counts(G, 0);
  auto worker = [&](int k) {
    wg_ctx *c = m->ctx[k];
    cudaSetDevice(c->device);
    // instance i -> device i mod G
    const long long Bk = instances > k ? (instances - k + G - 1) / G : 0;
    counts[k] = Bk;
    if (Bk > 0x7fffffffLL) { rcs[k] = WG_ERR_INVALID; return; }
    std::vector<double> v((size_t)3 * Bk);
    for (long long j = 0; j < Bk; ++j) {
      const double *s = vel_ref + 3 * (size_t)(k + j * G);
      v[3 * j] = s[0]; v[3 * j + 1] = s[1]; v[3 * j + 2] = s[2];
    }
    void *st = nullptr, *dv = nullptr;
    int rc = wg_malloc_device(c, sizeof(wg_herdt_mpc_state) * (size_t)std::max<long long>(Bk, 1), &st);
    if (rc == WG_OK) rc = wg_malloc_device(c, sizeof(double) * 3 * (size_t)std::max<long long>(Bk, 1), &dv);
    if (rc == WG_OK && Bk) rc = wg_herdt_mpc_init(c, WG_MEM_DEVICE, (int)Bk, init9, 0, static_cast<wg_herdt_mpc_state *>(st));
    if (rc == WG_OK && Bk) rc = wg_memcpy_h2d(c, dv, v.data(), sizeof(double) * 3 * (size_t)Bk);
    if (rc == WG_OK) rc = wg_sync(c);
    d_states[k] = st;
    if (rc == WG_OK) {
      wg_launch_count_reset(c);
      wg_timer_start(c);
      for (int done = 0; done < periods && rc == WG_OK && Bk; done += chunk) {
        const int nn = std::min(chunk, periods - done);
        rc = wg_herdt_mpc_run_batch(c, WG_MEM_DEVICE, (int)Bk, nn, static_cast<wg_herdt_mpc_state *>(st),
                                    done == 0 ? static_cast<const double *>(dv) : nullptr, nullptr, nullptr, nullptr);
      }
      wg_timer_stop_ms(c, &ms[k]);
      launches[k] = wg_launch_count(c);
    }
    if (rc == WG_OK && m->d_stats[k]) {
      rc = wg_herdt_mpc_stats(c, (int)Bk, static_cast<const wg_herdt_mpc_state *>(st), m->d_stats[k]);
      const double t = ms[k];
      if (rc == WG_OK) rc = wg_memcpy_h2d(c, m->d_stats[k] + 4, &t, sizeof(double));
      if (rc == WG_OK) rc = wg_sync(c);
    }
    if (dv) wg_free_device(c, dv);
    rcs[k] = rc;
  };
  {
    std::vector<std::thread> th;
    for (int k = 0; k < G; ++k) th.emplace_back(worker, k);
    for (auto &t : th) t.join();
  }
  int rc = WG_OK;
  for (int k = 0; k < G; ++k) if (rcs[k] != WG_OK) { rc = rcs[k]; snprintf(m->err, sizeof m->err, "device %d: %s", m->dev[k], wg_last_error(m->ctx[k])); }
  // ---- the one collective: SUM of the statistics, MAX of the time, over the devices
  double h[8] = {0};
  if (rc == WG_OK) {
    bool reduced = false;
    if (m->use_nccl) {
      bool ok = m->nccl.GroupStart() == 0;
      for (int k = 0; k < G && ok; ++k)
        ok = m->nccl.AllReduce(m->d_stats[k], m->d_stats[k], 4, NCCL_FLOAT64, NCCL_SUM, m->comm[k],
                               static_cast<cudaStream_t>(wg_ctx_stream(m->ctx[k]))) == 0;
      for (int k = 0; k < G && ok; ++k)
        ok = m->nccl.AllReduce(m->d_stats[k] + 4, m->d_stats[k] + 4, 1, NCCL_FLOAT64, NCCL_MAX, m->comm[k],
                               static_cast<cudaStream_t>(wg_ctx_stream(m->ctx[k]))) == 0;
      ok = (m->nccl.GroupEnd() == 0) && ok;
      for (int k = 0; k < G; ++k) wg_sync(m->ctx[k]);
      if (ok) {
        cudaSetDevice(m->dev[0]);
        reduced = wg_memcpy_d2h(m->ctx[0], h, m->d_stats[0], sizeof(double) * 8) == WG_OK && wg_sync(m->ctx[0]) == WG_OK;
      }
      if (!reduced) snprintf(m->err, sizeof m->err, "NCCL all-reduce failed: statistics are reduced on the host");
    }
    out->reduced_by_nccl = reduced ? 1 : 0;
    if (!reduced) {
      for (int k = 0; k < G; ++k) {
        double hk[8] = {0};
        cudaSetDevice(m->dev[k]);
        if (m->d_stats[k] && wg_memcpy_d2h(m->ctx[k], hk, m->d_stats[k], sizeof(double) * 8) == WG_OK && wg_sync(m->ctx[k]) == WG_OK) {
          for (int i = 0; i < 4; ++i) h[i] += hk[i];
          h[4] = std::max(h[4], hk[4]);
        }
      }
    }
  }
  for (int k = 0; k < G; ++k) if (d_states[k]) { cudaSetDevice(m->dev[k]); wg_free_device(m->ctx[k], d_states[k]); }
  out->instances = instances; out->periods = periods;
  out->qp_solves = h[0]; out->failures = h[1]; out->iterations = h[2]; out->still_online = h[3];
  out->seconds = h[4] * 1e-3;
  for (int k = 0; k < G && k < 16; ++k) { out->device_ms[k] = ms[k]; out->device_instances[k] = counts[k]; out->device_launches[k] = launches[k]; }
  out->nccl_version = m->use_nccl ? m->nccl_version : 0;
  return rc;
}

}  // extern "C"
