// Masked depthwise conv1d over padded bf16 rows [B, C, pitch].
//
//   y[b, c, t'] = sum_k w[c, k] * xm[b, c, t'*S - P + k*D],   xm = x with frames t >= len_in[b] zeroed
//   y[b, c, t'] = 0 for t' >= len_out[b]   (the mask the following pointwise MaskedConv1d applies to its input)
//
// Replaces MaskedConv1d.forward for groups == channels (src/thunder/quartznet/blocks.py:158-182, built at
// quartznet/blocks.py:195-201).  There is no BatchNorm/ReLU between the depthwise and the pointwise conv
// (quartznet/blocks.py:193-224), so this kernel is a pure masked FIR; fp32 accumulation, bf16 I/O.
//
// Fast path (template on K, S, D with "same" padding): one warp produces 256 consecutive outputs of one
// (b, c) row; each lane owns 8 consecutive outputs and slides a register window over the taps
// (8 FMAs per weight, 32 FMAs per 128-bit shared-memory load).  The input span is staged once in shared memory
// as fp32 with 4 floats of padding per 32 so that the lane-strided 128-bit loads are bank-conflict free.
// Generic path: any K / stride / dilation / padding, one thread per output (small layers, tests).
#include "ts_common.cuh"

namespace ts {
namespace dw {

constexpr int WARPS = 8;
constexpr int U = 8;            // outputs per lane
constexpr int SEG = 32 * U;     // outputs per warp work item

__host__ __device__ constexpr int same_pad(int K, int S, int D) { return D > 1 ? (D * (K - 1) + 1) / 2 : K / 2; }
__host__ __device__ constexpr int phys(int j) { return j + 4 * (j >> 5); }

__device__ __forceinline__ int floor_div(int a, int b) {  // b > 0
  int q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}
__device__ __forceinline__ int out_len(int len, int K, int S, int D, int P) {
  return floor_div(len + 2 * P - D * (K - 1) - 1, S) + 1;
}

template <int K, int S, int D>
struct Geo {
  static constexpr int P = same_pad(K, S, D);
  static constexpr int OFF = ((-P) % 8 + 8) % 8;                       // logical index of the first needed input
  static constexpr int KP = (K + 3) / 4 * 4;                           // taps padded to a multiple of 4
  static constexpr int WIN = OFF + (U - 1) * S + (K - 1) * D + 1;      // per-lane window (logical floats)
  static constexpr int NV = (WIN + 3) / 4;                             // 128-bit loads per lane
  static constexpr int LIN = ((31 * U * S + NV * 4) + 7) / 8 * 8;      // staged logical floats per warp
  static constexpr int XS = phys(LIN) + 4;                             // physical floats per warp
  static constexpr int SMEM_PER_WARP = XS + KP;
};

template <int K, int S, int D>
__global__ void __launch_bounds__(WARPS * 32)
dw_fast_kernel(const __nv_bfloat16* __restrict__ x, int C, int T_in, int pitch_in, const float* __restrict__ w,
               const int32_t* __restrict__ len_in, __nv_bfloat16* __restrict__ y, int T_out, int pitch_out, int segs,
               int items, int f16) {
  using G = Geo<K, S, D>;
  extern __shared__ __align__(16) float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * WARPS + warp;
  if (item >= items) return;
  float* xs = smem_f + warp * G::SMEM_PER_WARP;
  float* ws = xs + G::XS;
  const int row = item / segs, seg = item - row * segs;
  const int b = row / C, c = row - b * C;
  const int t0 = seg * SEG;                  // first output of this warp
  const int a0 = t0 * S - G::P - G::OFF;     // input frame of logical index 0 (multiple of 8)
  int lin = T_in;
  if (len_in != nullptr) lin = min(lin, max(len_in[b], 0));
  const int lout = (len_in != nullptr) ? min(T_out, max(out_len(len_in[b], K, S, D, G::P), 0)) : T_out;

  // ---- stage weights (warp-uniform channel) and the masked input span -------------------------------
  for (int k = lane; k < G::KP; k += 32) ws[k] = (k < K) ? w[(size_t)c * K + k] : 0.f;
  const __nv_bfloat16* xrow = x + (size_t)row * pitch_in;
  for (int j = lane * 8; j < G::LIN; j += 256) {
    const int t = a0 + j;  // multiple of 8
    float v[8];
    if (t >= 0 && t + 8 <= pitch_in && t < lin) {
      const uint4 u = *reinterpret_cast<const uint4*>(xrow + t);
      const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 f = unpack16x2(q[h], f16 != 0);
        v[2 * h] = (t + 2 * h < lin) ? f.x : 0.f;
        v[2 * h + 1] = (t + 2 * h + 1 < lin) ? f.y : 0.f;
      }
    } else {
#pragma unroll
      for (int h = 0; h < 8; ++h) v[h] = 0.f;
    }
    float4* d = reinterpret_cast<float4*>(xs + phys(j));  // j multiple of 8: both halves stay inside a 32-block
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncwarp();

  // ---- register-sliding FIR ------------------------------------------------------------------------
  const int jb = lane * U * S;
  float xw[G::NV * 4];
  float acc[U];
#pragma unroll
  for (int u = 0; u < U; ++u) acc[u] = 0.f;
  float wk[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < K; ++k) {
    // just-in-time 128-bit window loads: tap k needs logical floats up to OFF + (U-1)*S + k*D
    const int need = (G::OFF + (U - 1) * S + k * D) / 4;
    const int have = (k == 0) ? -1 : (G::OFF + (U - 1) * S + (k - 1) * D) / 4;
#pragma unroll
    for (int i = 0; i < G::NV; ++i) {
      if (i > have && i <= need) {
        const float4 f = *reinterpret_cast<const float4*>(xs + phys(jb + 4 * i));
        xw[4 * i] = f.x; xw[4 * i + 1] = f.y; xw[4 * i + 2] = f.z; xw[4 * i + 3] = f.w;
      }
    }
    if ((k & 3) == 0) {
      const float4 w4 = *reinterpret_cast<const float4*>(ws + k);
      wk[0] = w4.x; wk[1] = w4.y; wk[2] = w4.z; wk[3] = w4.w;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = fmaf(wk[k & 3], xw[G::OFF + u * S + k * D], acc[u]);
  }
  // ---- masked bf16 store: 16 bytes per lane, 512 contiguous bytes per warp --------------------------
  const int to = t0 + lane * U;
  if (to < pitch_out) {
    uint32_t o[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float lo = (to + 2 * h < lout) ? acc[2 * h] : 0.f;
      const float hi = (to + 2 * h + 1 < lout) ? acc[2 * h + 1] : 0.f;
      o[h] = pack16x2(lo, hi, f16 != 0);
    }
    *reinterpret_cast<uint4*>(y + (size_t)row * pitch_out + to) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// one thread per output frame; rows x pitch_out threads
__global__ void dw_generic_kernel(const __nv_bfloat16* __restrict__ x, int C, int T_in, int pitch_in,
                                  const float* __restrict__ w, int K, int S, int D, int P,
                                  const int32_t* __restrict__ len_in, __nv_bfloat16* __restrict__ y, int T_out,
                                  int pitch_out, int rows, int f16) {
  const int t = blockIdx.y * blockDim.x + threadIdx.x;
  const int row = blockIdx.x;
  if (t >= pitch_out || row >= rows) return;
  const int b = row / C, c = row - b * C;
  int lin = T_in;
  if (len_in != nullptr) lin = min(lin, max(len_in[b], 0));
  const int lout = (len_in != nullptr) ? min(T_out, max(out_len(len_in[b], K, S, D, P), 0)) : T_out;
  float acc = 0.f;
  if (t < lout) {
    const __nv_bfloat16* xrow = x + (size_t)row * pitch_in;
    const float* wr = w + (size_t)c * K;
    for (int k = 0; k < K; ++k) {
      const int ti = t * S - P + k * D;
      if (ti >= 0 && ti < lin)
        acc = fmaf(wr[k], unpack16(reinterpret_cast<const uint16_t*>(xrow)[ti], f16 != 0), acc);
    }
  }
  reinterpret_cast<uint16_t*>(y)[(size_t)row * pitch_out + t] = pack16(acc, f16 != 0);
}

template <int K, int S, int D>
int launch_fast(const __nv_bfloat16* x, int B, int C, int T_in, int pitch_in, const float* w, const int32_t* len_in,
                __nv_bfloat16* y, int T_out, int pitch_out, cudaStream_t st, int f16) {
  using G = Geo<K, S, D>;
  const int segs = ceil_div(pitch_out, SEG);
  const long long items_ll = (long long)B * C * segs;
  TS_REQUIRE(items_ll < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_dw_conv: too many work items");
  const int items = (int)items_ll;
  const size_t smem = (size_t)WARPS * G::SMEM_PER_WARP * sizeof(float);
  auto kern = dw_fast_kernel<K, S, D>;
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    TS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  kern<<<ceil_div(items, WARPS), WARPS * 32, smem, st>>>(x, C, T_in, pitch_in, w, len_in, y, T_out, pitch_out, segs,
                                                         items, f16);
  TS_LAUNCH_CHECK("dw_fast_kernel");
  return TS_OK;
}

}  // namespace dw
}  // namespace ts

namespace ts {
int launch_dw_mma(const __nv_bfloat16* x, int B, int C, int T, int pitch_in, const float* w, int K, int P,
                  const int32_t* lens, __nv_bfloat16* y, int pitch_out, cudaStream_t st);
int launch_dw_tma(const __nv_bfloat16* x, int B, int C, int T, int pitch_in, const float* w, int K, int D, int P,
                  const int32_t* lens, __nv_bfloat16* y, int pitch_out, cudaStream_t st, int f16);
int launch_dw_persist(const __nv_bfloat16* x, int B, int C, int T, int pitch_in, const float* w, int K, int D, int P,
                      const int32_t* lens, __nv_bfloat16* y, int pitch_out, cudaStream_t st, int f16);
int option_dw_persist();
int option_dw_persist_c();
int small_footprint(long long frames);
}
using namespace ts;

#define TS_DW_CASE(KK, SS, DD)                                                                            \
  if (K == KK && S == SS && D == DD && P == dw::same_pad(KK, SS, DD))                                     \
    return dw::launch_fast<KK, SS, DD>(xb, B, C, T_in, pitch_in, w, len_in, yb, T_out, pitch_out, st, f16);

extern "C" int ts_dw_conv(const void* x, int B, int C, int T_in, int pitch_in, const float* w, int K, int S, int D,
                          int P, const int32_t* len_in, int flags, void* y, int pitch_out, void* stream) {
  TS_REQUIRE(x && w && y, TS_ERR_INVALID, "ts_dw_conv: null pointer");
  TS_REQUIRE(B > 0 && C > 0 && T_in > 0 && K > 0 && S > 0 && D > 0 && P >= 0, TS_ERR_INVALID, "ts_dw_conv: bad sizes");
  TS_REQUIRE(!(S > 1 && D > 1), TS_ERR_INVALID, "Only stride OR dilation may be greater than 1");
  const int T_out = (T_in + 2 * P - D * (K - 1) - 1) / S + 1;
  TS_REQUIRE(T_in + 2 * P - D * (K - 1) - 1 >= 0 && T_out > 0, TS_ERR_INVALID, "ts_dw_conv: empty output (T_in=%d K=%d)",
             T_in, K);
  TS_REQUIRE(pitch_in % 8 == 0 && pitch_in >= T_in, TS_ERR_INVALID, "ts_dw_conv: pitch_in must be a multiple of 8, >= T_in");
  TS_REQUIRE(pitch_out % 8 == 0 && pitch_out >= T_out, TS_ERR_INVALID,
             "ts_dw_conv: pitch_out must be a multiple of 8 and >= T_out=%d (got %d)", T_out, pitch_out);
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(y);
  cudaStream_t st = (cudaStream_t)stream;
  const int f16 = (flags & TS_ROWS_F16) ? 1 : 0;
  // stride-1 / dilation-1 / odd-K "same" convolutions: Toeplitz MMA on the tensor cores (dwmma.cu)
  // ... through TMA when the caller guarantees rows are already zero beyond len_in (or there are no lengths)
  if (S == 1 && option_dw_mma() && option_dw_tma() && (len_in == nullptr || (flags & TS_DW_INPUT_PREMASKED))) {
    // Persistent multi-channel CTAs (dwmma3.cu) where one CTA per (channel, tile group) cannot amortise its prologue: few
    // tiles per channel AND more than two waves of channels (Citrinet's 1024-channel layers at small per-GPU batches / short
    // sequences), and the 5-block Toeplitz set of the dilated layer.  Measured on B200 inside the full models
    // (tools/ab_opts.sh "dw_persist=1" "dw_persist=0", per-channel kernel with 4 input stages): Citrinet-1024 at 16 / 32 x 20 s
    // persistent 5.01 / 7.40 ms vs per-channel 5.48 / 7.80 ms; QuartzNet 15x5 (256 / 512 channels: one or two waves of
    // per-channel CTAs) at 16 / 32 / 64 x 15 s per-channel 1.63 / 2.26 / 3.41 ms vs persistent 1.77 / 2.30 / 3.55 ms, and at
    // 256 x 15 s (29 tiles per channel) 11.85 vs 12.25 ms.  Option dw_persist: 0 = never, 1 = this heuristic, 2 = always.
    {
      const int W = pitch_in / 64, HL = ceil_div(P, 64), NQ = HL + 1 + (63 + P) / 64;
      const int HR = NQ - 1 - HL, R = W + (HL > HR ? HL : HR);
      const int tiles_per_chan = R > 0 && R <= 128 ? ceil_div(B, 128 / R) : 1 << 30;
      if (option_dw_persist() >= 2 ||
          (option_dw_persist() == 1 &&
           (NQ > 3 || (tiles_per_chan <= 8 && C > option_dw_persist_c()) || small_footprint((long long)B * pitch_in)))) {
        const int rc = launch_dw_persist(xb, B, C, T_in, pitch_in, w, K, D, P, len_in, yb, pitch_out, st, f16);
        if (rc != TS_ERR_UNSUPPORTED) return rc;
      }
    }
    const int rc = launch_dw_tma(xb, B, C, T_in, pitch_in, w, K, D, P, len_in, yb, pitch_out, st, f16);
    if (rc != TS_ERR_UNSUPPORTED) return rc;
  }
  if (S == 1 && D == 1 && option_dw_mma() && !f16) {   // (the cp.async predecessor is bf16 only)
    const int rc = launch_dw_mma(xb, B, C, T_in, pitch_in, w, K, P, len_in, yb, pitch_out, st);
    if (rc != TS_ERR_UNSUPPORTED) return rc;
  }
  // QuartzNet (quartznet/blocks.py:341-410)
  TS_DW_CASE(33, 2, 1) TS_DW_CASE(33, 1, 1) TS_DW_CASE(39, 1, 1) TS_DW_CASE(51, 1, 1) TS_DW_CASE(63, 1, 1)
  TS_DW_CASE(75, 1, 1) TS_DW_CASE(87, 1, 2)
  // Citrinet-1024
  TS_DW_CASE(5, 1, 1) TS_DW_CASE(11, 1, 1) TS_DW_CASE(13, 1, 1) TS_DW_CASE(15, 1, 1) TS_DW_CASE(17, 1, 1)
  TS_DW_CASE(19, 1, 1) TS_DW_CASE(21, 1, 1) TS_DW_CASE(23, 1, 1) TS_DW_CASE(25, 1, 1) TS_DW_CASE(27, 1, 1)
  TS_DW_CASE(29, 1, 1) TS_DW_CASE(31, 1, 1) TS_DW_CASE(35, 1, 1) TS_DW_CASE(37, 1, 1) TS_DW_CASE(41, 1, 1)
  TS_DW_CASE(11, 2, 1) TS_DW_CASE(13, 2, 1) TS_DW_CASE(25, 2, 1)
  // generic
  const int rows = B * C;
  dim3 grid(rows, ceil_div(pitch_out, 128));
  dw::dw_generic_kernel<<<grid, 128, 0, st>>>(xb, C, T_in, pitch_in, w, K, S, D, P, len_in, yb, T_out, pitch_out, rows,
                                              f16);
  TS_LAUNCH_CHECK("dw_generic_kernel");
  return TS_OK;
}
