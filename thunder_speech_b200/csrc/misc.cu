// Small kernels around the two hot ones: layout conversion at module boundaries, per-layer length
// bookkeeping, the SqueezeExcite FC micro-kernel and the greedy CTC argmax + collapse.
#include "ts_common.cuh"

namespace ts {
namespace misc {

__device__ __forceinline__ float ld_as_float(const float* p) { return *p; }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float ld_as_float(const __half* p) { return __half2float(*p); }

// [B, C, T] (contiguous, f32 or bf16) -> bf16 rows [B, C, pitch]; frames t >= min(T, lens[b]) are zero-filled.
template <typename InT>
__global__ void pack_rows_kernel(const InT* __restrict__ in, int T, uint16_t* __restrict__ out, int pitch,
                                 long long rows, int C, const int32_t* __restrict__ lens, int f16) {
  const long long row = blockIdx.x;
  const int t = blockIdx.y * blockDim.x + threadIdx.x;
  if (row >= rows || t >= pitch) return;
  int lim = T;
  if (lens != nullptr) lim = min(lim, lens[row / C]);  // MaskedConv1d.mask_fill of the consumer (quartznet/blocks.py:158-167)
  const float v = (t < lim) ? ld_as_float(in + row * T + t) : 0.f;
  out[row * pitch + t] = pack16(v, f16 != 0);
}

// 16-bit rows [B, C, pitch] -> contiguous f32 [B, C, T]
__global__ void unpack_rows_kernel(const uint16_t* __restrict__ in, int pitch, float* __restrict__ out, int T,
                                   long long rows, int f16) {
  const long long row = blockIdx.x;
  const int t = blockIdx.y * blockDim.x + threadIdx.x;
  if (row >= rows || t >= T) return;
  out[row * T + t] = unpack16(in[row * pitch + t], f16 != 0);
}

// out[b] = clamp-free MaskedConv1d.get_seq_len (quartznet/blocks.py:142-156): floor((L + 2p - d(k-1) - 1)/s) + 1
__global__ void conv_lengths_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int B, int K, int S,
                                    int D, int P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int num = in[b] + 2 * P - D * (K - 1) - 1;
  int q = num / S;
  if (num % S != 0 && num < 0) --q;
  out[b] = q + 1;
}

__global__ void lengths_i64_to_i32_kernel(const int64_t* __restrict__ in, int32_t* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int64_t v = in[b];
  v = v > 2147483647ll ? 2147483647ll : (v < -2147483648ll ? -2147483648ll : v);
  out[b] = (int32_t)v;
}
__global__ void lengths_i32_to_i64_kernel(const int32_t* __restrict__ in, int64_t* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = in[b];
}

// SqueezeExcite excitation (citrinet/blocks.py:77-83): gate[b, c] = sigmoid(W2 relu(W1 (pool[b, :] / T))).
// One CTA per batch element; pool holds the per-channel SUMS over all T frames (no mask -- the reference pools
// with AdaptiveAvgPool1d over the whole padded time axis).
// Two kernels so that both FCs spread over (units / 16) x B CTAs instead of one CTA per utterance:
//   se_hidden_kernel: hid[b, h] = relu(sum_c W1[h, c] * pool[b, c] / T)      grid (ceil(H/16), B), 8 warps x 2 units
//   se_gate_kernel:   gate[b, c] = sigmoid(sum_h W2[c, h] * hid[b, h])       grid (ceil(C/64), B), 8 warps x 8 units
template <bool FIXED>
__global__ void __launch_bounds__(256)
se_hidden_kernel(const void* __restrict__ pool_, float inv_T, const float* __restrict__ w1, int C, int H,
                 float* __restrict__ hid) {
  extern __shared__ float sm[];   // mean[C]
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const size_t i = (size_t)b * C + c;
    const float v = FIXED ? se_pool_value(static_cast<const long long*>(pool_)[i]) : static_cast<const float*>(pool_)[i];
    sm[c] = v * inv_T;
  }
  __syncthreads();
  const int h0 = blockIdx.x * 16 + warp * 2;
  float a0 = 0.f, a1 = 0.f;
  const bool ok0 = h0 < H, ok1 = h0 + 1 < H;
  const float* r0 = w1 + (size_t)h0 * C;
  const float* r1 = r0 + C;
#pragma unroll 4
  for (int c = lane; c < C; c += 32) {
    const float mv = sm[c];
    if (ok0) a0 = fmaf(__ldg(r0 + c), mv, a0);
    if (ok1) a1 = fmaf(__ldg(r1 + c), mv, a1);
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if (lane == 0) {
    if (ok0) hid[(size_t)b * H + h0] = fmaxf(a0, 0.f);
    if (ok1) hid[(size_t)b * H + h0 + 1] = fmaxf(a1, 0.f);
  }
}

__global__ void __launch_bounds__(256)
se_gate_kernel(const float* __restrict__ hid, const float* __restrict__ w2, int C, int H, float* __restrict__ gate) {
  extern __shared__ float sm[];   // hid[H]
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = threadIdx.x; h < H; h += blockDim.x) sm[h] = hid[(size_t)b * H + h];
  __syncthreads();
  const int c0 = blockIdx.x * 64 + warp * 8;
  float a[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) a[u] = 0.f;
  for (int h = lane; h < H; h += 32) {
    const float hv = sm[h];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (c0 + u < C) a[u] = fmaf(__ldg(w2 + (size_t)(c0 + u) * H + h), hv, a[u]);
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float r = warp_sum(a[u]);
    if (lane == 0 && c0 + u < C) gate[(size_t)b * C + c0 + u] = 1.f / (1.f + __expf(-r));
  }
}

// Strided 1x1 "conv" gather for residual branches with stride (citrinet/blocks.py:159-168): out[b, c, t'] =
// x[b, c, S t'] for S t' < min(T_in, len_in[b]), else 0 (MaskedConv1d.mask_fill); pad frames zero.  8 outputs/thread.
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ x, int C, int T_in, int pitch_in, int S,
                                   const int32_t* __restrict__ len_in, __nv_bfloat16* __restrict__ y, int pitch_out,
                                   long long rows) {
  const long long row = blockIdx.x;
  const int t0 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (row >= rows || t0 >= pitch_out) return;
  int lin = T_in;
  if (len_in != nullptr) lin = min(lin, max(len_in[row / C], 0));
  const __nv_bfloat16* xr = x + row * pitch_in;
  uint32_t o[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const int ta = (t0 + 2 * h) * S, tb = (t0 + 2 * h + 1) * S;
    const unsigned short a = (ta < lin) ? *reinterpret_cast<const unsigned short*>(xr + ta) : (unsigned short)0;
    const unsigned short b = (tb < lin) ? *reinterpret_cast<const unsigned short*>(xr + tb) : (unsigned short)0;
    o[h] = (uint32_t)a | ((uint32_t)b << 16);
  }
  *reinterpret_cast<uint4*>(y + row * pitch_out + t0) = make_uint4(o[0], o[1], o[2], o[3]);
}

// im2col over 16-bit rows: y[b, c * K + k, t] = x[b, c, t * S + k * D - P] inside [0, min(T_in, len_in[b])), else 0; zero in
// [T_out, pitch_out).  Turns a NON-separable MaskedConv1d with kernel_size > 1 (the reference's QuartznetBlock default,
// src/thunder/quartznet/blocks.py:212-219) into ONE pointwise GEMM with Cin * K input channels on the pair-GEMM kernel --
// conv.weight [Cout, Cin, K] is that GEMM's [Cout, Cin * K] weight as it lies in memory.  8 frames per thread.
__global__ void im2col_rows_kernel(const unsigned short* __restrict__ x, int C, int T_in, int pitch_in, int K, int S, int D,
                                   int P, const int32_t* __restrict__ len_in, unsigned short* __restrict__ y, int T_out,
                                   int pitch_out, long long rows, int rows_out) {
  const long long row = blockIdx.x;            // (b, c, k)
  const int t0 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (row >= rows || t0 >= pitch_out) return;
  const int k = (int)(row % K);
  const long long bc = row / K;
  int lin = T_in;
  if (len_in != nullptr) lin = min(lin, max(len_in[bc / C], 0));
  const unsigned short* xr = x + bc * pitch_in;
  uint32_t o[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    unsigned short v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int t = t0 + 2 * h + e;
      const int src = t * S + k * D - P;
      v[e] = (t < T_out && src >= 0 && src < lin) ? xr[src] : (unsigned short)0;
    }
    o[h] = (uint32_t)v[0] | ((uint32_t)v[1] << 16);
  }
  // utterance b's rows start at b * rows_out (>= C * K: the GEMM's K extent may be padded by the caller)
  const long long b = bc / C, ck = row - b * C * K;
  *reinterpret_cast<uint4*>(y + (b * rows_out + ck) * pitch_out + t0) = make_uint4(o[0], o[1], o[2], o[3]);
}

// SE excite for blocks WITHOUT a residual branch (Citrinet stem / epilogue, citrinet/blocks.py:154,195-197):
// out = relu(gate[b, c] * y1[b, c, t]), frames t >= lens[b] stored as zero when lens is given.  8 frames per thread.
__global__ void se_apply_kernel(const __nv_bfloat16* __restrict__ y1, const float* __restrict__ gate, int C, int pitch,
                                const int32_t* __restrict__ lens, int relu, __nv_bfloat16* __restrict__ out,
                                long long rows, int f16) {
  const long long row = blockIdx.x;
  const int t = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (row >= rows || t >= pitch) return;
  const float g = gate[row];
  const int lim = lens ? lens[row / C] : pitch;
  const uint4 u = *reinterpret_cast<const uint4*>(y1 + row * pitch + t);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  uint32_t o[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float2 f = unpack16x2(w[h], f16 != 0);
    float lo = g * f.x, hi = g * f.y;
    if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
    if (t + 2 * h >= lim) lo = 0.f;
    if (t + 2 * h + 1 >= lim) hi = 0.f;
    o[h] = pack16x2(lo, hi, f16 != 0);
  }
  *reinterpret_cast<uint4*>(out + row * pitch + t) = make_uint4(o[0], o[1], o[2], o[3]);
}

// Greedy CTC decode (module.py:100 `pred.argmax(1)`; text_processing/transform.py:107-110 unique_consecutive).
// One CTA per utterance.  Phase 1: argmax over the vocabulary axis of logits[b, :, t] (first maximal index wins,
// NaN counts as maximal -- torch.argmax semantics).  Phase 2: warp 0 collapses consecutive repeats with ballot
// + popcount prefix sums.  Blanks are kept (the reference strips the blank *string* after joining) unless
// drop_blank >= 0.
template <typename InT>
__global__ void __launch_bounds__(128)
ctc_argmax_kernel(const InT* __restrict__ logits, int V, int T, int pitch, int64_t* __restrict__ ids) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const InT* base = logits + (size_t)b * V * pitch + t;
  float best = ld_as_float(base);
  int bi = 0;
  bool best_nan = best != best;
  int v = 1;
  for (; v + 4 <= V && !best_nan; v += 4) {   // 4 independent coalesced loads in flight
    float x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = ld_as_float(base + (size_t)(v + u) * pitch);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (best_nan) break;
      if (x[u] != x[u]) {
        bi = v + u;
        best_nan = true;
      } else if (x[u] > best) {
        best = x[u];
        bi = v + u;
      }
    }
  }
  for (; v < V && !best_nan; ++v) {
    const float x = ld_as_float(base + (size_t)v * pitch);
    if (x != x) {
      bi = v;
      best_nan = true;
    } else if (x > best) {
      best = x;
      bi = v;
    }
  }
  ids[(size_t)b * T + t] = bi;
}

// one warp per utterance: collapse consecutive repeats with ballot + popcount prefix sums
__global__ void __launch_bounds__(32)
ctc_collapse_kernel(const int64_t* __restrict__ ids, int T, int64_t* __restrict__ collapsed,
                    int32_t* __restrict__ counts, int drop_blank) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int64_t* row = ids + (size_t)b * T;
  int n = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    bool keep = false;
    int64_t id = -1;
    if (t < T) {
      id = row[t];
      keep = (t == 0) || (id != row[t - 1]);
      if (drop_blank >= 0 && id == drop_blank) keep = false;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) collapsed[(size_t)b * T + n + __popc(m & ((1u << lane) - 1u))] = id;
    n += __popc(m);
  }
  for (int t = n + lane; t < T; t += 32) collapsed[(size_t)b * T + t] = -1;
  if (lane == 0) counts[b] = n;
}

}  // namespace misc
}  // namespace ts

using namespace ts;

extern "C" int ts_pack_rows(const void* in, int in_dtype, int B, int C, int T, const int32_t* lens, void* out,
                            int out_dtype, int pitch, void* stream) {
  TS_REQUIRE(in && out, TS_ERR_INVALID, "ts_pack_rows: null pointer");
  TS_REQUIRE(out_dtype == TS_BF16 || out_dtype == TS_F16, TS_ERR_INVALID, "ts_pack_rows: rows are TS_BF16 or TS_F16");
  const int f16 = out_dtype == TS_F16;
  TS_REQUIRE(B > 0 && C > 0 && T > 0 && pitch >= T, TS_ERR_INVALID, "ts_pack_rows: bad sizes");
  const long long rows = (long long)B * C;
  TS_REQUIRE(rows < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_pack_rows: too many rows");
  dim3 grid((unsigned)rows, ceil_div(pitch, 256));
  if (in_dtype == TS_F32)
    misc::pack_rows_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)in, T, (uint16_t*)out, pitch, rows, C, lens, f16);
  else if (in_dtype == TS_BF16)
    misc::pack_rows_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, T,
                                                                                  (uint16_t*)out, pitch, rows, C, lens, f16);
  else if (in_dtype == TS_F16)
    misc::pack_rows_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)in, T, (uint16_t*)out, pitch,
                                                                           rows, C, lens, f16);
  else
    TS_REQUIRE(false, TS_ERR_INVALID, "ts_pack_rows: bad input dtype %d", in_dtype);
  TS_LAUNCH_CHECK("pack_rows_kernel");
  return TS_OK;
}

extern "C" int ts_unpack_rows(const void* in, int in_dtype, int pitch, int B, int C, int T, float* out, void* stream) {
  TS_REQUIRE(in && out, TS_ERR_INVALID, "ts_unpack_rows: null pointer");
  TS_REQUIRE(in_dtype == TS_BF16 || in_dtype == TS_F16, TS_ERR_INVALID, "ts_unpack_rows: rows are TS_BF16 or TS_F16");
  TS_REQUIRE(B > 0 && C > 0 && T > 0 && pitch >= T, TS_ERR_INVALID, "ts_unpack_rows: bad sizes");
  const long long rows = (long long)B * C;
  TS_REQUIRE(rows < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_unpack_rows: too many rows");
  dim3 grid((unsigned)rows, ceil_div(T, 256));
  misc::unpack_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)in, pitch, out, T, rows,
                                                                   in_dtype == TS_F16);
  TS_LAUNCH_CHECK("unpack_rows_kernel");
  return TS_OK;
}

extern "C" int ts_conv_lengths(const int32_t* in, int32_t* out, int B, int K, int S, int D, int P, void* stream) {
  TS_REQUIRE(in && out && B > 0 && K > 0 && S > 0 && D > 0 && P >= 0, TS_ERR_INVALID, "ts_conv_lengths: bad arguments");
  misc::conv_lengths_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(in, out, B, K, S, D, P);
  TS_LAUNCH_CHECK("conv_lengths_kernel");
  return TS_OK;
}

extern "C" int ts_lengths_to_i32(const int64_t* in, int32_t* out, int B, void* stream) {
  TS_REQUIRE(in && out && B > 0, TS_ERR_INVALID, "ts_lengths_to_i32: bad arguments");
  misc::lengths_i64_to_i32_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(in, out, B);
  TS_LAUNCH_CHECK("lengths_i64_to_i32_kernel");
  return TS_OK;
}

extern "C" int ts_lengths_to_i64(const int32_t* in, int64_t* out, int B, void* stream) {
  TS_REQUIRE(in && out && B > 0, TS_ERR_INVALID, "ts_lengths_to_i64: bad arguments");
  misc::lengths_i32_to_i64_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(in, out, B);
  TS_LAUNCH_CHECK("lengths_i32_to_i64_kernel");
  return TS_OK;
}

extern "C" int ts_se_fc(const void* pool, int pool_dtype, int B, int C, int H, int T, const float* w1, const float* w2,
                        float* hid, float* gate, void* stream) {
  TS_REQUIRE(pool_dtype == TS_F32 || pool_dtype == TS_FIX32, TS_ERR_INVALID, "ts_se_fc: pool must be TS_F32 or TS_FIX32");
  TS_REQUIRE(pool && w1 && w2 && hid && gate, TS_ERR_INVALID, "ts_se_fc: null pointer");
  TS_REQUIRE(B > 0 && C > 0 && H > 0 && T > 0 && B <= 65535, TS_ERR_INVALID, "ts_se_fc: bad sizes");
  TS_REQUIRE((size_t)C * sizeof(float) <= 48 * 1024 && (size_t)H * sizeof(float) <= 48 * 1024, TS_ERR_UNSUPPORTED,
             "ts_se_fc: C=%d / H=%d too large", C, H);
  dim3 g1(ceil_div(H, 16), B), g2(ceil_div(C, 64), B);
  if (pool_dtype == TS_FIX32)
    misc::se_hidden_kernel<true><<<g1, 256, C * sizeof(float), (cudaStream_t)stream>>>(pool, 1.0f / (float)T, w1, C, H, hid);
  else
    misc::se_hidden_kernel<false><<<g1, 256, C * sizeof(float), (cudaStream_t)stream>>>(pool, 1.0f / (float)T, w1, C, H, hid);
  TS_LAUNCH_CHECK("se_hidden_kernel");
  misc::se_gate_kernel<<<g2, 256, H * sizeof(float), (cudaStream_t)stream>>>(hid, w2, C, H, gate);
  TS_LAUNCH_CHECK("se_gate_kernel");
  return TS_OK;
}

namespace ts {
namespace misc {
// Large vocabularies (Citrinet: V = 1025): one thread per frame walking all V rows is latency bound (256 dependent rounds of
// four loads).  Here a CTA owns 32 frames and its 8 warps each take the rows v = w, w + 8, ... (32 consecutive frames per
// load: coalesced), then the 8 partial results per frame are merged with torch.argmax's rules: the FIRST NaN wins, otherwise
// the largest value with the SMALLEST index.
template <typename InT>
__global__ void __launch_bounds__(256)
ctc_argmax_wide_kernel(const InT* __restrict__ logits, int V, int T, int pitch, int64_t* __restrict__ ids) {
  __shared__ float s_best[8][32];
  __shared__ int s_idx[8][32];
  __shared__ int s_nan[8][32];
  const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + lane;
  float best = -INFINITY;
  int bi = -1, nan_i = INT_MAX;
  if (t < T) {
    const InT* base = logits + (size_t)b * V * pitch + t;
    int v = w;
    for (; v + 24 < V; v += 32) {   // 4 independent loads in flight per thread, 32 per frame across the warps
      float x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = ld_as_float(base + (size_t)(v + 8 * u) * pitch);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (x[u] != x[u]) nan_i = min(nan_i, v + 8 * u);
        else if (x[u] > best || bi < 0) { best = x[u]; bi = v + 8 * u; }
      }
    }
    for (; v < V; v += 8) {
      const float x = ld_as_float(base + (size_t)v * pitch);
      if (x != x) nan_i = min(nan_i, v);
      else if (x > best || bi < 0) { best = x; bi = v; }
    }
  }
  s_best[w][lane] = best;
  s_idx[w][lane] = bi;
  s_nan[w][lane] = nan_i;
  __syncthreads();
  if (w == 0 && t < T) {
    float gb = s_best[0][lane];
    int gi = s_idx[0][lane], gn = s_nan[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float vb = s_best[k][lane];
      const int vi = s_idx[k][lane];
      gn = min(gn, s_nan[k][lane]);
      if (vi >= 0 && (gi < 0 || vb > gb || (vb == gb && vi < gi))) { gb = vb; gi = vi; }
    }
    ids[(size_t)b * T + t] = gn != INT_MAX ? gn : gi;
  }
}
}  // namespace misc
}  // namespace ts

extern "C" int ts_ctc_greedy(const void* logits, int dtype, int B, int V, int T, int pitch, int64_t* ids,
                             int64_t* collapsed, int32_t* counts, int drop_blank, void* stream) {
  TS_REQUIRE(logits && ids && collapsed && counts, TS_ERR_INVALID, "ts_ctc_greedy: null pointer");
  TS_REQUIRE(B > 0 && V > 0 && T > 0 && pitch >= T && B <= 65535, TS_ERR_INVALID, "ts_ctc_greedy: bad sizes");
  dim3 grid(ceil_div(T, 128), B);
  if (V >= 128 && (dtype == TS_F32 || dtype == TS_BF16)) {   // word-piece vocabularies: rows split over the warps of a CTA
    dim3 gw(ceil_div(T, 32), B);
    if (dtype == TS_F32)
      misc::ctc_argmax_wide_kernel<float><<<gw, 256, 0, (cudaStream_t)stream>>>((const float*)logits, V, T, pitch, ids);
    else
      misc::ctc_argmax_wide_kernel<__nv_bfloat16><<<gw, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)logits, V, T, pitch,
                                                                                   ids);
  } else if (dtype == TS_F32) {
    misc::ctc_argmax_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>((const float*)logits, V, T, pitch, ids);
  } else if (dtype == TS_BF16) {
    misc::ctc_argmax_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)logits, V, T,
                                                                                 pitch, ids);
  } else {
    TS_REQUIRE(false, TS_ERR_INVALID, "ts_ctc_greedy: bad dtype %d", dtype);
  }
  TS_LAUNCH_CHECK("ctc_argmax_kernel");
  misc::ctc_collapse_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(ids, T, collapsed, counts, drop_blank);
  TS_LAUNCH_CHECK("ctc_collapse_kernel");
  return TS_OK;
}

extern "C" int ts_gather_rows(const void* x, int B, int C, int T_in, int pitch_in, int S, const int32_t* len_in,
                              void* y, int pitch_out, void* stream) {
  TS_REQUIRE(x && y, TS_ERR_INVALID, "ts_gather_rows: null pointer");
  TS_REQUIRE(B > 0 && C > 0 && T_in > 0 && S > 0 && pitch_in >= T_in && pitch_out % 8 == 0, TS_ERR_INVALID,
             "ts_gather_rows: bad sizes");
  const int T_out = (T_in - 1) / S + 1;
  TS_REQUIRE(pitch_out >= T_out, TS_ERR_INVALID, "ts_gather_rows: pitch_out %d < T_out %d", pitch_out, T_out);
  const long long rows = (long long)B * C;
  TS_REQUIRE(rows < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_gather_rows: too many rows");
  dim3 grid((unsigned)rows, ceil_div(pitch_out / 8, 128));
  misc::gather_rows_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, C, T_in, pitch_in, S, len_in,
                                                                   (__nv_bfloat16*)y, pitch_out, rows);
  TS_LAUNCH_CHECK("gather_rows_kernel");
  return TS_OK;
}

extern "C" int ts_im2col_rows(const void* x, int B, int C, int T_in, int pitch_in, int K, int S, int D, int P,
                              const int32_t* len_in, void* y, int rows_out, int pitch_out, void* stream) {
  TS_REQUIRE(x && y, TS_ERR_INVALID, "ts_im2col_rows: null pointer");
  TS_REQUIRE(B > 0 && C > 0 && T_in > 0 && K > 0 && S > 0 && D > 0 && P >= 0 && pitch_in >= T_in && pitch_out % 8 == 0,
             TS_ERR_INVALID, "ts_im2col_rows: bad sizes");
  const int T_out = (T_in + 2 * P - D * (K - 1) - 1) / S + 1;
  TS_REQUIRE(T_in + 2 * P - D * (K - 1) - 1 >= 0 && T_out > 0, TS_ERR_INVALID, "ts_im2col_rows: empty output (T_in=%d K=%d)",
             T_in, K);
  TS_REQUIRE(pitch_out >= T_out, TS_ERR_INVALID, "ts_im2col_rows: pitch_out %d < T_out %d", pitch_out, T_out);
  TS_REQUIRE(rows_out >= C * K, TS_ERR_INVALID, "ts_im2col_rows: rows_out %d < C * K = %d", rows_out, C * K);
  const long long rows = (long long)B * C * K;
  TS_REQUIRE(rows < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_im2col_rows: too many rows");
  dim3 grid((unsigned)rows, ceil_div(pitch_out / 8, 128));
  misc::im2col_rows_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const unsigned short*)x, C, T_in, pitch_in, K, S, D, P,
                                                                   len_in, (unsigned short*)y, T_out, pitch_out, rows,
                                                                   rows_out);
  TS_LAUNCH_CHECK("im2col_rows_kernel");
  return TS_OK;
}

extern "C" int ts_se_apply(const void* y1, const float* gate, int B, int C, int pitch, const int32_t* lens, int flags,
                           void* out, void* stream) {
  const int relu = flags & TS_PW_RELU;
  TS_REQUIRE(y1 && gate && out, TS_ERR_INVALID, "ts_se_apply: null pointer");
  TS_REQUIRE(B > 0 && C > 0 && pitch > 0 && pitch % 8 == 0, TS_ERR_INVALID, "ts_se_apply: bad sizes");
  const long long rows = (long long)B * C;
  TS_REQUIRE(rows < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_se_apply: too many rows");
  dim3 grid((unsigned)rows, ceil_div(pitch / 8, 128));
  misc::se_apply_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)y1, gate, C, pitch, lens, relu,
                                                                (__nv_bfloat16*)out, rows, (flags & TS_ROWS_F16) ? 1 : 0);
  TS_LAUNCH_CHECK("se_apply_kernel");
  return TS_OK;
}

// ---- SpecAugment / SpecCutout (training-time feature masking, src/thunder/quartznet/spec_augment.py) ---------------
// Zeroes, in place, every element (b, f, t) that lies inside ANY of the n rectangles [f0, f1) x [t0, t1); the same
// rectangles for every utterance (mask_along_axis: "all examples will have the same mask interval").  The rectangles live
// in DEVICE memory so that a CUDA-graph replay can use fresh draws.  One thread per element; only elements inside a
// rectangle are written.
namespace ts {
namespace misc {
template <typename T>
__global__ void spec_mask_kernel(T* __restrict__ feat, int C, int Tn, int pitch, const int32_t* __restrict__ rects, int n,
                                 long long rows) {
  extern __shared__ int32_t sr[];
  for (int i = threadIdx.x; i < 4 * n; i += blockDim.x) sr[i] = rects[i];
  __syncthreads();
  const long long row = blockIdx.x;
  const int t = blockIdx.y * blockDim.x + threadIdx.x;
  if (row >= rows || t >= Tn) return;
  const int f = (int)(row % C);
  bool hit = false;
  for (int i = 0; i < n; ++i) hit = hit || (f >= sr[4 * i] && f < sr[4 * i + 1] && t >= sr[4 * i + 2] && t < sr[4 * i + 3]);
  if (hit) feat[row * pitch + t] = T(0.f);
}
}  // namespace misc
}  // namespace ts

extern "C" int ts_spec_mask(void* feat, int dtype, int B, int C, int T, int pitch, const int32_t* rects, int n,
                            void* stream) {
  TS_REQUIRE(feat && (rects || n == 0), TS_ERR_INVALID, "ts_spec_mask: null pointer");
  TS_REQUIRE(dtype == TS_F32 || dtype == TS_BF16, TS_ERR_INVALID, "ts_spec_mask: bad dtype");
  TS_REQUIRE(B > 0 && C > 0 && T > 0 && pitch >= T && n >= 0 && n <= 1024, TS_ERR_INVALID, "ts_spec_mask: bad sizes");
  if (n == 0) return TS_OK;
  const long long rows = (long long)B * C;
  TS_REQUIRE(rows < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_spec_mask: too many rows");
  dim3 grid((unsigned)rows, ceil_div(T, 256));
  const size_t smem = (size_t)4 * n * sizeof(int32_t);
  if (dtype == TS_F32)
    misc::spec_mask_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>((float*)feat, C, T, pitch, rects, n, rows);
  else
    misc::spec_mask_kernel<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>((__nv_bfloat16*)feat, C, T, pitch, rects,
                                                                                      n, rows);
  TS_LAUNCH_CHECK("spec_mask_kernel");
  return TS_OK;
}

// ---- scatter_rows: the transpose of gather_rows (training backward of strided layers) --------------------------------
// dst[b, c, t] = (t % S == 0 && t / S < T_src) ? src[b, c, t / S] : 0   for t < T_dst (zero up to the pitch), plus the old
// dst value when `accumulate`.  Zero-upsampling turns a strided depthwise conv's input gradient into a stride-1 conv with
// flipped taps; with accumulate it adds a strided 1x1 residual conv's input gradient onto the main-branch gradient.
namespace ts {
namespace misc {
__global__ void scatter_rows_kernel(const __nv_bfloat16* __restrict__ src, int T_src, int pitch_src, int S,
                                    __nv_bfloat16* __restrict__ dst, int T_dst, int pitch_dst, int accumulate,
                                    long long rows) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row = blockIdx.x;
  const int t0 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (row >= rows || t0 >= pitch_dst) return;
  const __nv_bfloat16* sr = src + row * pitch_src;
  __nv_bfloat16* dr = dst + row * pitch_dst + t0;
  uint4 old = make_uint4(0, 0, 0, 0);
  if (accumulate) old = *reinterpret_cast<const uint4*>(dr);
  const uint32_t ow[4] = {old.x, old.y, old.z, old.w};
  uint32_t o[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int t = t0 + 2 * h + e;
      float val = 0.f;
      if (t < T_dst) {
        if (accumulate) val = __uint_as_float(e == 0 ? (ow[h] << 16) : (ow[h] & 0xFFFF0000u));
        const int q = t / S;
        if (q * S == t && q < T_src) val += __bfloat162float(sr[q]);
      }
      v[e] = val;
    }
    __nv_bfloat162 pr = __floats2bfloat162_rn(v[0], v[1]);
    o[h] = *reinterpret_cast<uint32_t*>(&pr);
  }
  *reinterpret_cast<uint4*>(dr) = make_uint4(o[0], o[1], o[2], o[3]);
}
}  // namespace misc
}  // namespace ts

extern "C" int ts_scatter_rows(const void* src, int B, int C, int T_src, int pitch_src, int S, void* dst, int T_dst,
                               int pitch_dst, int accumulate, void* stream) {
  TS_REQUIRE(src && dst, TS_ERR_INVALID, "ts_scatter_rows: null pointer");
  TS_REQUIRE(B > 0 && C > 0 && T_src > 0 && T_dst > 0 && S > 0 && pitch_src >= T_src && pitch_dst >= T_dst &&
                 pitch_dst % 8 == 0,
             TS_ERR_INVALID, "ts_scatter_rows: bad sizes");
  TS_REQUIRE((T_src - 1) * S < T_dst, TS_ERR_INVALID, "ts_scatter_rows: source does not fit: (T_src-1)*S >= T_dst");
  const long long rows = (long long)B * C;
  TS_REQUIRE(rows < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_scatter_rows: too many rows");
  dim3 grid((unsigned)rows, ceil_div(pitch_dst / 8, 128));
  TS_CUDA(launch_pdl(misc::scatter_rows_kernel, grid, dim3(128), 0, (cudaStream_t)stream, (option_pdl() & 2) != 0,
                     (const __nv_bfloat16*)src, T_src, pitch_src, S, (__nv_bfloat16*)dst, T_dst, pitch_dst, accumulate, rows));
  TS_LAUNCH_CHECK("scatter_rows_kernel");
  return TS_OK;
}
