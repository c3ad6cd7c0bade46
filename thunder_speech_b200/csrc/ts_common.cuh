// Shared host/device helpers for libthunder_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <utility>

#include "../../include/thunder_b200.h"

namespace ts {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int option_pw_big();  // 1: bf16-row GEMMs with Cout > 128 use the persistent 256x256-tile kernel (pwgemm2.cu)
int option_dw_tma();  // 1: pre-masked stride-1 depthwise convs use the TMA Toeplitz kernel (dwmma2.cu)
int option_dw_mma();  // 1: stride-1 depthwise convs run on the tensor cores (dwmma.cu), 0: SIMT kernels only
int option_pw_bn();   // experiment switch (see api.cu)
int option_serpentine();  // 1: consecutive launches walk the utterances in alternating directions (L2 reuse)
int option_dbg();     // scratch knob for experiments
// L2 reuse between consecutive layers: the two streaming kernels (Toeplitz conv, pair GEMM) each move 2-4x the 126 MB L2
// per launch, so when a kernel ends only the LAST ~60-100 MB it wrote are still resident.  Every launch of those kernels
// therefore walks the utterances in the direction opposite to its predecessor's -- it starts on what is still in L2.
// Pure scheduling hint (results are order independent); returns 0 / 1 alternately, always 0 with option serpentine = 0.
int next_walk_reversed();

#define TS_REQUIRE(cond, code, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      ::ts::set_error(__VA_ARGS__);            \
      return (code);                           \
    }                                          \
  } while (0)

#define TS_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::ts::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                      __LINE__);                                                        \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

// after a kernel launch: pick up launch-configuration errors without synchronising
#define TS_LAUNCH_CHECK(name)                                                      \
  do {                                                                             \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) {                                                       \
      ::ts::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));    \
      return (int)_e;                                                              \
    }                                                                              \
    ::ts::count_launch();                                                          \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

constexpr int kRowPitchAlign = 64;  // frames; 128 bytes of bf16

int option_pdl();  // bit 0: the hot inference kernels (pair GEMM, Toeplitz conv) launch with programmatic stream serialization (PDL); bit 1: the training-step kernels too

// ---- programmatic dependent launch --------------------------------------------------------------
// Device side: `pdl_launch_dependents()` lets the NEXT kernel in the stream start its prologue as soon as every CTA of
// this grid has executed it; `pdl_wait()` blocks until the PREVIOUS grid has completed and its writes are visible.
// Both are no-ops when the kernel was launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Host side: launch with the programmatic-stream-serialization attribute (captured into CUDA graphs as a
// programmatic dependency edge).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- phase tracing (diagnostics; ts_trace; compiled in only with -DTS_TRACE=1: `make trace` -> libthunder_b200_trace.so) ---
// When a trace buffer is installed, every launch of the hot kernels takes the next 32-word slot; CTA 0 stamps
// %globaltimer (ns) into words 1..9 and the last CTA into words 17..25 at fixed points of its life (entry, prologue done,
// predecessor complete, first operands, last MMA issued, first accumulator, last store issued, stores drained, exit);
// words 10..15 of CTA 0 = cycles its role threads spent waiting (operands, free accumulator, free stage, accumulator
// ready, staging tile read by the previous TMA store) and the life of the MMA issue loop.
// Word 0 = kernel id (1 pair GEMM, 2 persistent Toeplitz, 3 per-channel Toeplitz) | grid << 8.  nullptr = tracing off.
// The production library compiles all of this to nothing (measured cost of the live hooks: 1-3 % on the GEMMs).
#ifndef TS_TRACE
#define TS_TRACE 0
#endif
unsigned long long* trace_next_slot(int kernel_id, unsigned grid);
#if TS_TRACE
__device__ __forceinline__ void trace_head(unsigned long long* slot, int which, int kernel_id) {
  if (slot != nullptr && which == 0) slot[0] = (unsigned long long)kernel_id | ((unsigned long long)gridDim.x << 8);
}
__device__ __forceinline__ void trace_stamp(unsigned long long* slot, int which, int i) {
  if (slot != nullptr && which >= 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    slot[which * 16 + i] = t;
  }
}
// stall accounting: cycles spent inside a wait are summed in a REGISTER of the waiting thread and written once
// (trace_put) when its role loop ends -- a global read-modify-write per wait would dominate what it measures
#define TS_TIMED_WAIT(on, acc, stmt)     \
  do {                                   \
    long long _t0 = 0;                   \
    if (on) _t0 = clock64();             \
    stmt;                                \
    if (on) (acc) += clock64() - _t0;    \
  } while (0)
// dbg bit 7 (128): every CTA also stamps its entry / exit time into words 32 + 2 * cta (+1) behind the slot (the reader
// must have allocated 32 + 2 * grid words: single-launch diagnostics only, tools/trace_dw.py)
__device__ __forceinline__ void trace_cta(unsigned long long* slot, int dbg, int cta, int end) {
  if (slot != nullptr && (dbg & 128)) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    slot[32 + 2 * cta + end] = t;
  }
}
__device__ __forceinline__ void trace_put(unsigned long long* slot, int which, int word, long long v) {
  if (slot != nullptr && which == 0) slot[word] = (unsigned long long)v;
}
__device__ __forceinline__ long long trace_clock() { return clock64(); }
#else
__device__ __forceinline__ void trace_head(unsigned long long*, int, int) {}
__device__ __forceinline__ void trace_stamp(unsigned long long*, int, int) {}
#define TS_TIMED_WAIT(on, acc, stmt) \
  do {                               \
    stmt;                            \
  } while (0)
__device__ __forceinline__ void trace_cta(unsigned long long*, int, int, int) {}
__device__ __forceinline__ void trace_put(unsigned long long*, int, int, long long) {}
__device__ __forceinline__ long long trace_clock() { return 0; }
#endif

// ---- device helpers ---------------------------------------------------------------------------
// SqueezeExcite time-pooling accumulator: partial sums are added as 64-bit FIXED-POINT integers (units of 2^-32), so the
// result does not depend on the order in which the tiles of an utterance arrive (fp32 atomics do: last-bit differences of
// the gate were amplified to O(1) logit differences, run to run, by a 115-layer random-init Citrinet).  |sum| < 2^31.
constexpr float kSePoolScale = 4294967296.f;   // 2^32
__device__ __forceinline__ void se_pool_add(unsigned long long* slot, float partial) {
  atomicAdd(slot, static_cast<unsigned long long>(__float2ll_rn(partial * kSePoolScale)));
}
__device__ __forceinline__ float se_pool_value(long long fixed) { return static_cast<float>(static_cast<double>(fixed) * (1.0 / 4294967296.0)); }

// ---- 16-bit row formats ------------------------------------------------------------------------
// Activation rows are bf16 (default) or IEEE fp16 ("half rows": 11-bit mantissa, the mode that holds the 2e-2 logit
// parity on the 15x5-deep networks; same bytes, same tcgen05 kind::f16 rate).  fp16 conversions SATURATE to +-65504
// instead of producing inf, so an out-of-range activation degrades gracefully.  `f16` is kernel-uniform.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// ... with the ReLU folded into the conversion (F2FP.RELU: one instruction instead of two FMNMX + one F2FP per pair)
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi, bool f16) {
  return f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
__device__ __forceinline__ float2 unpack16x2(uint32_t w, bool f16) { return f16 ? unpack_f16x2(w) : unpack_bf16x2(w); }
__device__ __forceinline__ float unpack16(uint16_t h, bool f16) {
  return f16 ? __half2float(__ushort_as_half(h)) : __uint_as_float((uint32_t)h << 16);
}
__device__ __forceinline__ uint16_t pack16(float v, bool f16) { return (uint16_t)(pack16x2(v, 0.f, f16) & 0xFFFFu); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ts
