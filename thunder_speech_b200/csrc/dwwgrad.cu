// Depthwise-conv tap gradient, register-tiled:
//
//     dw[c, k] = sum_b sum_t da[b, c, t] * x[b, c, t + k*D - P]          (stride 1, D = 1 or 2)
//
// One CTA owns one channel and a chunk of R utterances.  The R rows of `x` (with a zero halo, masked to the utterance
// length) and of `da` (masked to T_out) are staged in shared memory as bf16.  Work is cut into tasks of 8 time steps x 8
// taps: a thread holds 8 tap accumulators in registers for its tap block kb = tid % NKB and walks over (row, time block)
// pairs; per task it issues 3 (D = 1) or 4 (D = 2) 16-byte shared loads and 64 FMAs.  The staging offset P8 = round_up(P, 8)
// makes every window start 16-byte aligned: x[t + k*D - P] = sx[t + (k + delta/D)*D] with delta = P8 - P, so the kernel
// simply computes taps k' = k + delta/D and the epilogue drops the shifted-out ones.  The per-thread partials are summed
// over the thread groups in a fixed order (deterministic), and each CTA writes part[chunk, c, :]; the caller sums chunks.
#include "ts_common.cuh"

namespace ts {
namespace dwwg {

constexpr int THREADS = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    v[2 * h] = __uint_as_float(w[h] << 16);
    v[2 * h + 1] = __uint_as_float(w[h] & 0xFFFF0000u);
  }
}

// zero the bf16 lanes of `u` at positions >= n (0 <= n <= 8)
__device__ __forceinline__ uint4 keep_first(uint4 u, int n) {
  uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    if (2 * h >= n) w[h] = 0u;
    else if (2 * h + 1 >= n) w[h] &= 0x0000FFFFu;
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int D>
__global__ void __launch_bounds__(THREADS)
dw_wgrad_tiled_kernel(const __nv_bfloat16* __restrict__ da, int T_out, int pitch_out, const __nv_bfloat16* __restrict__ x,
                      int T_in, int pitch_in, const int32_t* __restrict__ len_in, int B, int C, int K, int P, int P8,
                      int R, int NKB, int XP, float* __restrict__ part) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sx = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [R][XP]
  __nv_bfloat16* sd = sx + (size_t)R * XP;                                 // [R][pitch_out]
  float* red = reinterpret_cast<float*>(sd + (size_t)R * pitch_out);       // [NG][NKB * 8]
  const int c = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
  const int b0 = chunk * R, nrow = min(R, B - b0);
  const int NTB = (T_out + 7) >> 3;

  // ---- stage rows: sx[r][m] = x[b0 + r, c, m - P8] (0 outside [0, len)),  sd[r][t] = da[b0 + r, c, t] (0 for t >= T_out)
  const int xchunks = XP >> 3, dchunks = pitch_out >> 3;
  for (int i = tid; i < nrow * xchunks; i += THREADS) {
    const int r = i / xchunks, m = (i - r * xchunks) << 3;
    const int t = m - P8;
    int lin = T_in;
    if (len_in != nullptr) lin = min(lin, max(__ldg(len_in + b0 + r), 0));
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (t >= 0 && t < lin) {
      u = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b0 + r) * C + c) * pitch_in + t));
      if (t + 8 > lin) u = keep_first(u, lin - t);
    }
    *reinterpret_cast<uint4*>(sx + (size_t)r * XP + m) = u;
  }
  for (int i = tid; i < nrow * dchunks; i += THREADS) {
    const int r = i / dchunks, t = (i - r * dchunks) << 3;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (t < T_out) {
      u = __ldg(reinterpret_cast<const uint4*>(da + ((size_t)(b0 + r) * C + c) * pitch_out + t));
      if (t + 8 > T_out) u = keep_first(u, T_out - t);
    }
    *reinterpret_cast<uint4*>(sd + (size_t)r * pitch_out + t) = u;
  }
  __syncthreads();

  // ---- tasks
  const int NG = THREADS / NKB;          // thread groups, each walks (row, time block) pairs with stride NG
  const int kb = tid % NKB, grp = tid / NKB;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (grp < NG) {
    const int npairs = nrow * NTB;
    int r = grp / NTB, tb = grp - r * NTB;
    const int dr = NG / NTB, dtb = NG - dr * NTB;
    for (int pidx = grp; pidx < npairs; pidx += NG) {
      float d[8];
      unpack8(*reinterpret_cast<const uint4*>(sd + (size_t)r * pitch_out + (tb << 3)), d);
      constexpr int NW = (8 + 7 * D + 7) / 8;   // 16-byte chunks of the x window: 2 (D = 1), 3 (D = 2)
      float xw[NW * 8];
      const __nv_bfloat16* xs = sx + (size_t)r * XP + (tb << 3) + kb * 8 * D;
#pragma unroll
      for (int w = 0; w < NW; ++w) unpack8(*reinterpret_cast<const uint4*>(xs + 8 * w), xw + 8 * w);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[kk] = fmaf(d[j], xw[j + kk * D], acc[kk]);
      r += dr;
      tb += dtb;
      if (tb >= NTB) {
        tb -= NTB;
        ++r;
      }
    }
  }
  if (grp < NG) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) red[(size_t)grp * (NKB * 8) + kb * 8 + kk] = acc[kk];
  }
  __syncthreads();
  // ---- deterministic sum over groups, drop the alignment shift
  const int shift = (P8 - P) / D;
  if (tid < K) {
    float s = 0.f;
    for (int g = 0; g < NG; ++g) s += red[(size_t)g * (NKB * 8) + tid + shift];
    part[((size_t)chunk * C + c) * K + tid] = s;
  }
}

}  // namespace dwwg
}  // namespace ts

using namespace ts;

// returns TS_ERR_UNSUPPORTED when the shape is outside this kernel's envelope (caller falls back)
int launch_dw_wgrad_tiled(const __nv_bfloat16* da, int T_out, int pitch_out, const __nv_bfloat16* x, int T_in, int pitch_in,
                          const int32_t* len_in, int B, int C, int K, int D, int P, int bchunk, float* part,
                          cudaStream_t st) {
  if (D != 1 && D != 2) return TS_ERR_UNSUPPORTED;
  const int P8 = round_up(P, 8);
  if ((P8 - P) % D != 0 || K > dwwg::THREADS) return TS_ERR_UNSUPPORTED;
  const int shift = (P8 - P) / D;
  const int NKB = ceil_div(K + shift, 8);
  if (NKB > 32) return TS_ERR_UNSUPPORTED;
  const int NTB = ceil_div(T_out, 8);
  // furthest staged element any task touches: window of the last time block and last tap block
  const int need = (NTB - 1) * 8 + (NKB - 1) * 8 * D + ((8 + 7 * D + 7) / 8) * 8;
  const int XP = round_up(need, 8);
  const int NG = dwwg::THREADS / NKB;
  const size_t smem = (size_t)bchunk * (XP + pitch_out) * 2 + (size_t)NG * NKB * 8 * 4;
  if (smem > 200 * 1024) return TS_ERR_UNSUPPORTED;
  dim3 grid(C, ceil_div(B, bchunk));
  if (D == 1) {
    if (smem > 48 * 1024)
      TS_CUDA(cudaFuncSetAttribute(dwwg::dw_wgrad_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dwwg::dw_wgrad_tiled_kernel<1><<<grid, dwwg::THREADS, smem, st>>>(da, T_out, pitch_out, x, T_in, pitch_in, len_in, B, C, K,
                                                                     P, P8, bchunk, NKB, XP, part);
  } else {
    if (smem > 48 * 1024)
      TS_CUDA(cudaFuncSetAttribute(dwwg::dw_wgrad_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dwwg::dw_wgrad_tiled_kernel<2><<<grid, dwwg::THREADS, smem, st>>>(da, T_out, pitch_out, x, T_in, pitch_in, len_in, B, C, K,
                                                                     P, P8, bchunk, NKB, XP, part);
  }
  TS_LAUNCH_CHECK("dw_wgrad_tiled_kernel");
  return TS_OK;
}
