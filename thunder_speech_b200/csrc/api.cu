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

// runtime switches (ts_set_option / THUNDER_B200_OPTIONS), one table so that adding an A/B knob is one line
struct Option { const char* name; std::atomic<int> value; };
static Option g_options[] = {
    {"dw_mma", {1}},          // stride-1 depthwise convs on the tensor cores
    {"pw_big", {1}},          // persistent 256-wide GEMM tiles for 16-bit outputs with Cout > 128
    {"dw_tma", {1}},          // TMA-fed Toeplitz kernel for pre-masked rows
    {"dw_persist", {1}},      // persistent multi-channel Toeplitz CTAs (dwmma3.cu): 0 never, 1 heuristic (few tiles per channel), 2 always
    {"pw_bn", {0}},
    {"pw_resident", {0}},     // pair GEMM keeps the weight rows of its m-tile in shared memory when K <= 512 (measured: -7 %, off)
    {"pdl", {3}},             // programmatic dependent launch: bit 0 inference kernels, bit 1 training kernels
    {"pw_pair", {2}},         // CTA-pair GEMM: 1 = only K >= 1024, 2 = every 16-bit-row GEMM with Cout > 128
    {"dw_base_offset", {0}},
    {"dw_share_halo", {1}},
    {"dw_pro", {30}},         // per-CTA prologue of the Toeplitz kernel in tenths of a tile (grid cost model)
    {"dw_persist_c", {592}},  // ... heuristic: persistent kernel only above this many channels (two waves of per-channel CTAs)
    {"dw_nstage", {0}},       // per-channel Toeplitz kernel: input stages (0 = default 4)
    {"tma_l2", {3}},          // L2 promotion of every tensor map: 0 none, 1 64 B, 2 128 B, 3 256 B
    {"dwp_nbuf", {0}},        // persistent Toeplitz kernel: Toeplitz buffers (0 = auto: 2 when NQ <= 3), experiment
    {"dwp_nstage", {0}},      // ... input stages (0 = auto)
    {"pw_ws", {0}},           // pair GEMM with the weights in tensor memory (pwgemm4.cu) for K <= 512
    {"small", {0}},           // small-footprint co-resident kernels: 0 never, 1 when B x pitch <= small_frames, 2 always
    {"small_frames", {32768}},
    {"serpentine", {1}},      // alternate the utterance walk direction between consecutive launches (L2 reuse)
    {"dbg", {0}},             // scratch knob for experiments
};
static std::atomic<int>& opt(const char* name) {
  for (auto& o : g_options)
    if (strcmp(o.name, name) == 0) return o.value;
  return g_options[sizeof(g_options) / sizeof(g_options[0]) - 1].value;
}
int option_pdl() { return opt("pdl").load(std::memory_order_relaxed); }
int option_pw_pair() { return opt("pw_pair").load(std::memory_order_relaxed); }
int option_pw_resident() { return opt("pw_resident").load(std::memory_order_relaxed); }
int option_pw_bn() { return opt("pw_bn").load(std::memory_order_relaxed); }
int option_dw_pro() { return opt("dw_pro").load(std::memory_order_relaxed); }
int option_dw_share_halo() { return opt("dw_share_halo").load(std::memory_order_relaxed); }
int option_dw_persist() { return opt("dw_persist").load(std::memory_order_relaxed); }
int option_dw_tma() { return opt("dw_tma").load(std::memory_order_relaxed); }
int option_dw_base_offset() { return opt("dw_base_offset").load(std::memory_order_relaxed); }
int option_dw_mma() { return opt("dw_mma").load(std::memory_order_relaxed); }
int option_pw_big() { return opt("pw_big").load(std::memory_order_relaxed); }
int small_footprint(long long frames) {
  const int m = opt("small").load(std::memory_order_relaxed);
  return m >= 2 || (m == 1 && frames <= (long long)opt("small_frames").load(std::memory_order_relaxed));
}
int option_dw_persist_c() { return opt("dw_persist_c").load(std::memory_order_relaxed); }
int option_dw_nstage() { return opt("dw_nstage").load(std::memory_order_relaxed); }
int option_tma_l2() { return opt("tma_l2").load(std::memory_order_relaxed); }
int option_dwp_nbuf() { return opt("dwp_nbuf").load(std::memory_order_relaxed); }
int option_dwp_nstage() { return opt("dwp_nstage").load(std::memory_order_relaxed); }
int option_pw_ws() { return opt("pw_ws").load(std::memory_order_relaxed); }
int option_serpentine() { return opt("serpentine").load(std::memory_order_relaxed); }
int option_dbg() { return opt("dbg").load(std::memory_order_relaxed); }
static std::atomic<unsigned> g_walk{0};
int next_walk_reversed() { return option_serpentine() ? (int)(g_walk.fetch_add(1, std::memory_order_relaxed) & 1u) : 0; }

static unsigned long long* g_trace = nullptr;
static std::atomic<int> g_trace_slots{0}, g_trace_next{0};
unsigned long long* trace_next_slot(int kernel_id, unsigned grid) {
  if (g_trace == nullptr) return nullptr;
  const int i = g_trace_next.fetch_add(1, std::memory_order_relaxed);
  if (i >= g_trace_slots.load(std::memory_order_relaxed)) return nullptr;
  (void)kernel_id; (void)grid;   // the header word is written by the traced CTA itself (launches may be under capture)
  return g_trace + (size_t)i * 32;
}

}  // namespace ts

extern "C" int ts_trace(void* buffer, int slots) {
#if !TS_TRACE
  if (buffer != nullptr) {
    ts::set_error("ts_trace: this build has no trace hooks (make -C thunder_speech_b200/csrc trace -> libthunder_b200_trace.so)");
    return TS_ERR_UNSUPPORTED;
  }
#endif
  ts::g_trace = reinterpret_cast<unsigned long long*>(buffer);
  ts::g_trace_slots.store(buffer ? slots : 0);
  ts::g_trace_next.store(0);
  return TS_OK;
}

extern "C" const char* ts_version(void) { return "thunder_b200 0.1 (sm_100a)"; }
extern "C" const char* ts_last_error(void) { return ts::g_err; }
extern "C" int64_t ts_launch_count(void) { return ts::g_launches.load(std::memory_order_relaxed); }
extern "C" int ts_row_pitch(int T) { return T <= 0 ? 0 : ts::round_up(T, ts::kRowPitchAlign); }

extern "C" int ts_set_option(const char* name, int value) {
  if (name != nullptr)
    for (auto& o : ts::g_options)
      if (strcmp(o.name, name) == 0) {
        o.value.store(value);
        return TS_OK;
      }
  ts::set_error("ts_set_option: unknown option '%s'", name ? name : "(null)");
  return TS_ERR_INVALID;
}
