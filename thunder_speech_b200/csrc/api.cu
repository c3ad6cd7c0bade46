// Library-level entry points of libthunder_b200.so (version, errors, launch accounting).
#include "ts_common.cuh"

#include <atomic>
#include <string.h>

namespace ts {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int> g_opt_dw_mma{1};
static std::atomic<int> g_opt_pw_big{1};
static std::atomic<int> g_opt_dw_tma{1};
static std::atomic<int> g_opt_pw_bn{0};
static std::atomic<int> g_opt_pdl{3};
int option_pdl() { return g_opt_pdl.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_pw_pair{2};
int option_pw_pair() { return g_opt_pw_pair.load(std::memory_order_relaxed); }
int option_pw_bn() { return g_opt_pw_bn.load(std::memory_order_relaxed); }
static std::atomic<int> g_opt_dw_base_offset{0};
static std::atomic<int> g_opt_dw_share_halo{1};
static std::atomic<int> g_opt_dw_pro{30};
int option_dw_pro() { return g_opt_dw_pro.load(std::memory_order_relaxed); }
int option_dw_share_halo() { return g_opt_dw_share_halo.load(std::memory_order_relaxed); }
int option_dw_tma() { return g_opt_dw_tma.load(std::memory_order_relaxed); }
int option_dw_base_offset() { return g_opt_dw_base_offset.load(std::memory_order_relaxed); }
int option_dw_mma() { return g_opt_dw_mma.load(std::memory_order_relaxed); }
int option_pw_big() { return g_opt_pw_big.load(std::memory_order_relaxed); }

}  // namespace ts

extern "C" const char* ts_version(void) { return "thunder_b200 0.1 (sm_100a)"; }
extern "C" const char* ts_last_error(void) { return ts::g_err; }
extern "C" int64_t ts_launch_count(void) { return ts::g_launches.load(std::memory_order_relaxed); }
extern "C" int ts_row_pitch(int T) { return T <= 0 ? 0 : ts::round_up(T, ts::kRowPitchAlign); }

extern "C" int ts_set_option(const char* name, int value) {
  if (name != nullptr && strcmp(name, "dw_mma") == 0) {
    ts::g_opt_dw_mma.store(value);
    return TS_OK;
  }
  if (name != nullptr && strcmp(name, "pdl") == 0) {
    ts::g_opt_pdl.store(value);
    return TS_OK;
  }
  if (name != nullptr && strcmp(name, "pw_pair") == 0) {
    ts::g_opt_pw_pair.store(value);
    return TS_OK;
  }
  if (name != nullptr && strcmp(name, "pw_bn") == 0) {
    ts::g_opt_pw_bn.store(value);
    return TS_OK;
  }
  if (name != nullptr && strcmp(name, "dw_tma") == 0) {
    ts::g_opt_dw_tma.store(value);
    return TS_OK;
  }
  if (name != nullptr && strcmp(name, "dw_base_offset") == 0) {
    ts::g_opt_dw_base_offset.store(value);
    return TS_OK;
  }
  if (name != nullptr && strcmp(name, "dw_pro") == 0) {
    ts::g_opt_dw_pro.store(value);
    return TS_OK;
  }
  if (name != nullptr && strcmp(name, "dw_share_halo") == 0) {
    ts::g_opt_dw_share_halo.store(value);
    return TS_OK;
  }
  if (name != nullptr && strcmp(name, "pw_big") == 0) {
    ts::g_opt_pw_big.store(value);
    return TS_OK;
  }
  ts::set_error("ts_set_option: unknown option '%s'", name ? name : "(null)");
  return TS_ERR_INVALID;
}
