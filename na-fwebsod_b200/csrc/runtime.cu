// Error string, version, tuning table.
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include "common.cuh"

namespace nawsod {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::mutex g_mu;
static std::map<std::string, int64_t>& tuning() { static std::map<std::string, int64_t> t; return t; }
int64_t get_tuning(const char* key, int64_t dflt) {
  std::lock_guard<std::mutex> l(g_mu);
  auto it = tuning().find(key);
  return it == tuning().end() ? dflt : it->second;
}
int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}
}  // namespace nawsod

extern "C" {
const char* nawsod_last_error(void) { return nawsod::g_err; }
int nawsod_version(void) { return 100; }
int nawsod_set_tuning(const char* key, int64_t value) {
  static const char* known[] = {"pool_slab_bytes", "pool_chunks", "pool_force_global", "pool_threads", "pool_generic", "pool_rowcache", "pool_rows2", "pool_skip_idle",
                                "gemm_pair", "gemm_tma_store", "gemm_max_ctas", "mil_ctas", "sgd_max_ctas", "p2p_ctas", nullptr};
  if (!key) { nawsod::set_error("nawsod_set_tuning: null key"); return NAWSOD_ERR_ARG; }
  for (int i = 0; known[i]; ++i)
    if (std::strcmp(known[i], key) == 0) {
      std::lock_guard<std::mutex> l(nawsod::g_mu);
      if (value < 0 && std::strcmp(key, "pool_rows2") == 0) nawsod::tuning().erase(key);   // back to the built-in default
      else nawsod::tuning()[key] = value;
      return NAWSOD_OK;
    }
  nawsod::set_error("nawsod_set_tuning: unknown key '%s'", key);
  return NAWSOD_ERR_ARG;
}
}
