// Library-level entry points: ABI version, thread-local error text, launch counter.
#include "common.cuh"
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace vr {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- tuning options ---------------------------------------------------------------------
static std::atomic<int> g_syrk_pair{[] { const char* e = getenv("VR_SYRK_PAIR"); return e ? atoi(e) : 1; }()};
static std::atomic<int> g_syrk_waves{[] { const char* e = getenv("VR_SYRK_WAVES"); return e ? atoi(e) : 1; }()};
static std::atomic<int> g_sweep_tile_a{8}, g_sweep_tile_b{8}, g_kendall_complement{1}, g_kendall_tree{1};
int option_kendall_complement() { return g_kendall_complement.load(std::memory_order_relaxed); }
int option_kendall_tree() { return g_kendall_tree.load(std::memory_order_relaxed); }
int option_sweep_tile_a() { return g_sweep_tile_a.load(std::memory_order_relaxed); }
int option_sweep_tile_b() { return g_sweep_tile_b.load(std::memory_order_relaxed); }
int option_syrk_pair() { return g_syrk_pair.load(std::memory_order_relaxed); }
int option_syrk_waves() { int w = g_syrk_waves.load(std::memory_order_relaxed); return w < 1 ? 1 : w; }

// ---- stage profiler ---------------------------------------------------------------------
static const char* kFamilyNames[F_NUM_FAMILIES] = {"masks", "presort", "count", "rank", "rank_fused", "pearson", "kendall",
                                            "rdm_center", "rdm_syrk", "rdm_syrk_pair", "rdm_epilogue", "row_rank", "column_corr"};
struct ProfRec { int fam; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;        // records of the current session
static std::vector<cudaEvent_t> g_pool;    // recycled events
static unsigned long long g_fam_launches[F_NUM_FAMILIES];
static thread_local int t_family = -1;

void count_launch(unsigned n) {
  g_launches.fetch_add(n, std::memory_order_relaxed);
  if (g_prof_on && t_family >= 0) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_fam_launches[t_family] += n;
  }
}

static cudaEvent_t take_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(int family, cudaStream_t s) : fam(family), stream(s), slot(-1) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof.size() >= (1u << 20)) return;
  ProfRec r{family, take_event(), take_event()};
  if (!r.a || !r.b) return;
  cudaEventRecord(r.a, stream);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
  t_family = family;
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].b, stream);
  t_family = -1;
}
}  // namespace vr

extern "C" int vr_profile_family_count(void) { return vr::F_NUM_FAMILIES; }
extern "C" const char* vr_profile_family_name(int i) { return (i >= 0 && i < vr::F_NUM_FAMILIES) ? vr::kFamilyNames[i] : ""; }
extern "C" int vr_profile_start(void) {
  std::lock_guard<std::mutex> lk(vr::g_prof_mu);
  for (auto& r : vr::g_prof) { vr::g_pool.push_back(r.a); vr::g_pool.push_back(r.b); }
  vr::g_prof.clear();
  for (int i = 0; i < vr::F_NUM_FAMILIES; ++i) vr::g_fam_launches[i] = 0;
  vr::g_prof_on = true;
  return VR_OK;
}
extern "C" int vr_profile_stop(double* ms, unsigned long long* scopes, unsigned long long* launches) {
  std::lock_guard<std::mutex> lk(vr::g_prof_mu);
  vr::g_prof_on = false;
  for (int i = 0; i < vr::F_NUM_FAMILIES; ++i) {
    if (ms) ms[i] = 0.0;
    if (scopes) scopes[i] = 0;
    if (launches) launches[i] = vr::g_fam_launches[i];
  }
  int rc = VR_OK;
  for (auto& r : vr::g_prof) {
    float t = 0.f;
    if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) {
      vr::set_error("vr_profile_stop: event query failed");
      rc = VR_ERR_CUDA;
      (void)cudaGetLastError();
    } else {
      if (ms) ms[r.fam] += (double)t;
      if (scopes) scopes[r.fam] += 1;
    }
    vr::g_pool.push_back(r.a);
    vr::g_pool.push_back(r.b);
  }
  vr::g_prof.clear();
  return rc;
}

extern "C" int vr_set_option(const char* name, int value) {
  VR_REQUIRE(name, "vr_set_option: null name");
  if (!strcmp(name, "syrk_pair")) { vr::g_syrk_pair.store(value); return VR_OK; }
  if (!strcmp(name, "syrk_waves")) { VR_REQUIRE(value >= 1 && value <= 64, "vr_set_option: syrk_waves=%d outside [1, 64]", value); vr::g_syrk_waves.store(value); return VR_OK; }
  if (!strcmp(name, "kendall_complement")) { vr::g_kendall_complement.store(value != 0); return VR_OK; }
  if (!strcmp(name, "kendall_tree")) {
    if (value < 0 || value > 2) { vr::set_error("vr_set_option: kendall_tree is 0 (off), 1 (when it pays) or 2 (always)"); return VR_ERR_INVALID; }
    vr::g_kendall_tree.store(value);
    return VR_OK;
  }
  if (!strcmp(name, "sweep_tile_a") || !strcmp(name, "sweep_tile_b")) {
    VR_REQUIRE(value >= 1 && value <= 8, "vr_set_option: %s=%d outside [1, 8]", name, value);
    (name[11] == 'a' ? vr::g_sweep_tile_a : vr::g_sweep_tile_b).store(value);
    return VR_OK;
  }
  vr::set_error("vr_set_option: unknown option '%s'", name);
  return VR_ERR_INVALID;
}
extern "C" int vr_get_option(const char* name, int* value) {
  VR_REQUIRE(name && value, "vr_get_option: null pointer");
  if (!strcmp(name, "syrk_pair")) { *value = vr::option_syrk_pair(); return VR_OK; }
  if (!strcmp(name, "syrk_waves")) { *value = vr::option_syrk_waves(); return VR_OK; }
  if (!strcmp(name, "kendall_complement")) { *value = vr::option_kendall_complement(); return VR_OK; }
  if (!strcmp(name, "kendall_tree")) { *value = vr::option_kendall_tree(); return VR_OK; }
  if (!strcmp(name, "sweep_tile_a")) { *value = vr::option_sweep_tile_a(); return VR_OK; }
  if (!strcmp(name, "sweep_tile_b")) { *value = vr::option_sweep_tile_b(); return VR_OK; }
  vr::set_error("vr_get_option: unknown option '%s'", name);
  return VR_ERR_INVALID;
}

extern "C" int vr_abi_version(void) { return VR_ABI_VERSION; }
extern "C" const char* vr_last_error_string(void) { return vr::g_err; }
extern "C" unsigned long long vr_launch_count(void) { return vr::g_launches.load(std::memory_order_relaxed); }
