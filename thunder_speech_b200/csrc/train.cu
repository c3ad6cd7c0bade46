// Training-step kernels around the two GEMM/Toeplitz engines (reference: BaseCTCModule.training_step,
// src/thunder/module.py:102-127, with train()-mode modules of src/thunder/quartznet/blocks.py).
//
// In train() mode nn.BatchNorm1d(eps=1e-3, momentum=0.1) (quartznet/blocks.py:222) normalises with BATCH statistics over
// all B x T positions (padded frames included -- the conv input is masked, the BatchNorm input is not), so BN can no
// longer be folded into the GEMM.  Per sub-block the forward is
//     a = dw(x)            (Toeplitz kernel)         z = W a        (pointwise GEMM, no shift / ReLU)
//     stats(z)             row_stats_kernel          y = relu(BN(z) [+ BN_r(z_r)])   bn_apply_kernel
// and the backward
//     dym = dy * (y > 0);  sums(dym, dym*z)          bn_bwd_reduce_kernel
//     dz = A dym + B z + C                           bn_bwd_apply_kernel
//     dW = sum_{b,t} dz a^T (wgrad GEMM, pwgemm_wgrad.cu);  da = W^T dz (pointwise GEMM);  dw taps: dw_wgrad_kernel;
//     dx = dw^T(da) (Toeplitz kernel with flipped taps).
// All tensors are bf16 rows [B, C, pitch]; reductions accumulate in fp32 and are written per (b, c) row (deterministic),
// the final sum over the batch is a tiny host-side torch reduction.
#include <algorithm>

#include "ts_common.cuh"

namespace ts {
namespace train {

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    v[2 * h] = __uint_as_float(w[h] << 16);
    v[2 * h + 1] = __uint_as_float(w[h] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint32_t o[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    __nv_bfloat162 pr = __floats2bfloat162_rn(v[2 * h], v[2 * h + 1]);
    o[h] = *reinterpret_cast<uint32_t*>(&pr);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

constexpr int RW = 8;  // warps per CTA for the row kernels

// stats[row] = (sum_t z, sum_t z^2) over t < T, one warp per (b, c) row
__global__ void __launch_bounds__(RW * 32)
row_stats_kernel(const __nv_bfloat16* __restrict__ z, int T, int pitch, long long rows, float2* __restrict__ stats) {
  pdl_launch_dependents();   // PDL: the next kernel may start its prologue; then wait for the previous grid
  pdl_wait();
  const long long row = (long long)blockIdx.x * RW + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* zr = z + row * pitch;
  float s = 0.f, ss = 0.f;
  for (int t = lane * 8; t < T; t += 256) {
    float v[8];
    unpack8(*reinterpret_cast<const uint4*>(zr + t), v);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (t + j < T) {
        s += v[j];
        ss = fmaf(v[j], v[j], ss);
      }
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (lane == 0) stats[row] = make_float2(s, ss);
}

// y = act(z * s[c] + h[c] (+ zr * sr[c] + hr[c])), frames t >= lens[b] (and the pad) stored as zero; 8 frames / thread
__global__ void bn_apply_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ s,
                                const float* __restrict__ h, const __nv_bfloat16* __restrict__ zr,
                                const float* __restrict__ sr, const float* __restrict__ hr, int C, int T, int pitch,
                                const int32_t* __restrict__ lens, int relu, __nv_bfloat16* __restrict__ y,
                                long long rows, const float* __restrict__ gate) {
  pdl_launch_dependents();   // PDL: the next kernel may start its prologue; then wait for the previous grid
  pdl_wait();
  const long long row = blockIdx.x;
  const int t = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (row >= rows || t >= pitch) return;
  const int c = (int)(row % C);
  const int lim = lens ? min(T, max(lens[row / C], 0)) : T;
  float v[8], o[8];
  unpack8(*reinterpret_cast<const uint4*>(z + row * pitch + t), v);
  const float sc = s[c], sh = h[c];
  const float gt = gate ? gate[row] : 1.f;   // SqueezeExcite: the main branch is scaled per (utterance, channel)
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = fmaf(v[j], sc, sh) * gt;
  if (zr != nullptr) {
    unpack8(*reinterpret_cast<const uint4*>(zr + row * pitch + t), v);
    const float sc2 = sr[c], sh2 = hr[c];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] += fmaf(v[j], sc2, sh2);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (relu) o[j] = fmaxf(o[j], 0.f);
    if (t + j >= lim) o[j] = 0.f;
  }
  *reinterpret_cast<uint4*>(y + row * pitch + t) = pack8(o);
}

// sums[row] = (sum dym, sum dym*z, sum dym*zr) over t < T with dym = dy * (y > 0) (relu) or dy
__global__ void __launch_bounds__(RW * 32)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                     const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ zr, int T, int pitch,
                     long long rows, int relu, float* __restrict__ sums, int C, const float* __restrict__ ms,
                     const float* __restrict__ mh, const float* __restrict__ msr, const float* __restrict__ mhr,
                     const float* __restrict__ gate) {
  pdl_launch_dependents();   // PDL: the next kernel may start its prologue; then wait for the previous grid
  pdl_wait();
  const long long row = (long long)blockIdx.x * RW + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  // y == nullptr: the ReLU mask is recomputed as (z * ms[c] + mh[c] (+ zr * msr[c] + mhr[c]) > 0), the expression
  // bn_apply_kernel evaluated in the forward pass -- saves reading y
  const bool remask = relu && y == nullptr;
  const int c = (int)(row % C);
  const float sc = remask ? ms[c] : 0.f, sh = remask ? mh[c] : 0.f;
  const float sc2 = (remask && zr != nullptr) ? msr[c] : 0.f, sh2 = (remask && zr != nullptr) ? mhr[c] : 0.f;
  const float gt = gate ? gate[row] : 1.f;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int t = lane * 8; t < T; t += 256) {
    float g[8], yy[8], zz[8], z2[8];
    unpack8(*reinterpret_cast<const uint4*>(dy + row * pitch + t), g);
    unpack8(*reinterpret_cast<const uint4*>(z + row * pitch + t), zz);
    if (relu && !remask) unpack8(*reinterpret_cast<const uint4*>(y + row * pitch + t), yy);
    if (zr != nullptr) unpack8(*reinterpret_cast<const uint4*>(zr + row * pitch + t), z2);
    if (remask) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float o = fmaf(zz[j], sc, sh) * gt;
        if (zr != nullptr) o += fmaf(z2[j], sc2, sh2);
        yy[j] = o;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (t + j < T) {
        const float d = (relu && !(yy[j] > 0.f)) ? 0.f : g[j];
        s0 += d;
        s1 = fmaf(d, zz[j], s1);
        if (zr != nullptr) s2 = fmaf(d, z2[j], s2);
      }
    }
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    sums[row * 3 + 0] = s0;
    sums[row * 3 + 1] = s1;
    sums[row * 3 + 2] = s2;
  }
}

// dz = A[c] dym + B[c] z + C[c]  (and dzr = Ar dym + Br zr + Cr) for t < T, pad frames zero
__global__ void bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                    const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ zr,
                                    const float* __restrict__ coef, const float* __restrict__ coef_r, int C, int T,
                                    int pitch, int relu, __nv_bfloat16* __restrict__ dz,
                                    __nv_bfloat16* __restrict__ dzr, long long rows, const float* __restrict__ ms,
                                    const float* __restrict__ mh, const float* __restrict__ msr,
                                    const float* __restrict__ mhr, const float* __restrict__ gate,
                                    const float* __restrict__ addc) {
  pdl_launch_dependents();   // PDL: the next kernel may start its prologue; then wait for the previous grid
  pdl_wait();
  const long long row = blockIdx.x;
  const int t = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (row >= rows || t >= pitch) return;
  const int c = (int)(row % C);
  float g[8], yy[8], zz[8], o[8];
  unpack8(*reinterpret_cast<const uint4*>(dy + row * pitch + t), g);
  unpack8(*reinterpret_cast<const uint4*>(z + row * pitch + t), zz);
  if (relu) {
    if (y != nullptr) {
      unpack8(*reinterpret_cast<const uint4*>(y + row * pitch + t), yy);
    } else {   // recompute the forward pre-activation exactly as bn_apply_kernel did
      const float sc = ms[c], sh = mh[c];
      const float gm = gate ? gate[row] : 1.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) yy[j] = fmaf(zz[j], sc, sh) * gm;
      if (zr != nullptr) {
        float z2[8];
        unpack8(*reinterpret_cast<const uint4*>(zr + row * pitch + t), z2);
        const float sc2 = msr[c], sh2 = mhr[c];
#pragma unroll
        for (int j = 0; j < 8; ++j) yy[j] += fmaf(z2[j], sc2, sh2);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(yy[j] > 0.f)) g[j] = 0.f;
  }
  const float a = coef[3 * c], b = coef[3 * c + 1], cc = coef[3 * c + 2];
  // SqueezeExcite: the gradient reaching the main branch's BatchNorm output is dym * gate[b,c] + addc[b,c] (the second
  // term is d loss / d mean_t(u), spread over every frame); the residual branch below sees dym unchanged
  const float gt = gate ? gate[row] : 1.f, ac = addc ? addc[row] : 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = (t + j < T) ? fmaf(a, fmaf(g[j], gt, ac), fmaf(b, zz[j], cc)) : 0.f;
  *reinterpret_cast<uint4*>(dz + row * pitch + t) = pack8(o);
  if (zr != nullptr) {
    unpack8(*reinterpret_cast<const uint4*>(zr + row * pitch + t), zz);
    const float a2 = coef_r[3 * c], b2 = coef_r[3 * c + 1], c2 = coef_r[3 * c + 2];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (t + j < T) ? fmaf(a2, g[j], fmaf(b2, zz[j], c2)) : 0.f;
    *reinterpret_cast<uint4*>(dzr + row * pitch + t) = pack8(o);
  }
}

// Depthwise weight gradient: part[bchunk, c, k] = sum_{b in chunk} sum_t' da[b, c, t'] * xm[b, c, t' S + k D - P]
// (xm = x masked at t >= len_in[b]; da is already zero beyond the output length).  One CTA per (channel, batch chunk);
// thread k owns tap k; both rows are staged in shared memory as fp32.
__global__ void __launch_bounds__(256)
dw_wgrad_kernel(const __nv_bfloat16* __restrict__ da, int T_out, int pitch_out, const __nv_bfloat16* __restrict__ x,
                int T_in, int pitch_in, const int32_t* __restrict__ len_in, int B, int C, int K, int S, int D, int P,
                int bchunk, float* __restrict__ part) {
  extern __shared__ float sm[];
  float* sx = sm;               // [pitch_in + 2 * halo] with halo = P on the left
  float* sd = sm + pitch_in + 2 * (K * D + 8);
  const int c = blockIdx.x, chunk = blockIdx.y;
  const int tid = threadIdx.x;
  const int halo = K * D + 8;
  float acc = 0.f;
  const int b0 = chunk * bchunk, b1 = min(B, b0 + bchunk);
  for (int b = b0; b < b1; ++b) {
    const long long row = (long long)b * C + c;
    int lin = T_in;
    if (len_in != nullptr) lin = min(lin, max(len_in[b], 0));
    __syncthreads();
    for (int i = tid; i < pitch_in + 2 * halo; i += blockDim.x) {
      const int t = i - halo;
      sx[i] = (t >= 0 && t < lin) ? __bfloat162float(x[row * pitch_in + t]) : 0.f;
    }
    for (int i = tid; i < pitch_out; i += blockDim.x)
      sd[i] = (i < T_out) ? __bfloat162float(da[row * pitch_out + i]) : 0.f;
    __syncthreads();
    if (tid < K) {
      const float* xs = sx + halo + tid * D - P;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int t = 0;
      for (; t + 4 <= T_out; t += 4) {
        a0 = fmaf(sd[t], xs[t * S], a0);
        a1 = fmaf(sd[t + 1], xs[(t + 1) * S], a1);
        a2 = fmaf(sd[t + 2], xs[(t + 2) * S], a2);
        a3 = fmaf(sd[t + 3], xs[(t + 3) * S], a3);
      }
      for (; t < T_out; ++t) a0 = fmaf(sd[t], xs[t * S], a0);
      acc += (a0 + a1) + (a2 + a3);
    }
  }
  if (tid < K) part[((size_t)chunk * C + c) * K + tid] = acc;
}

}  // namespace train
}  // namespace ts

using namespace ts;

extern "C" int ts_row_stats(const void* z, int B, int C, int T, int pitch, float* stats, void* stream) {
  TS_REQUIRE(z && stats, TS_ERR_INVALID, "ts_row_stats: null pointer");
  TS_REQUIRE(B > 0 && C > 0 && T > 0 && pitch >= T && pitch % 8 == 0, TS_ERR_INVALID, "ts_row_stats: bad sizes");
  const long long rows = (long long)B * C;
  TS_CUDA(launch_pdl(train::row_stats_kernel, dim3((unsigned)ceil_div64(rows, train::RW)), dim3(train::RW * 32), 0,
                     (cudaStream_t)stream, (option_pdl() & 2) != 0, (const __nv_bfloat16*)z, T, pitch, rows, (float2*)stats));
  TS_LAUNCH_CHECK("row_stats_kernel");
  return TS_OK;
}

extern "C" int ts_bn_apply(const void* z, const float* scale, const float* shift, const void* zr, const float* scale_r,
                           const float* shift_r, int B, int C, int T, int pitch, const int32_t* lens, int relu, void* y,
                           void* stream) {
  TS_REQUIRE(z && scale && shift && y, TS_ERR_INVALID, "ts_bn_apply: null pointer");
  TS_REQUIRE((zr == nullptr) == (scale_r == nullptr) && (zr == nullptr) == (shift_r == nullptr), TS_ERR_INVALID,
             "ts_bn_apply: residual operands disagree");
  TS_REQUIRE(B > 0 && C > 0 && T > 0 && pitch >= T && pitch % 8 == 0, TS_ERR_INVALID, "ts_bn_apply: bad sizes");
  const long long rows = (long long)B * C;
  TS_REQUIRE(rows < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_bn_apply: too many rows");
  dim3 grid((unsigned)rows, ceil_div(pitch / 8, 128));
  TS_CUDA(launch_pdl(train::bn_apply_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, (option_pdl() & 2) != 0,
                     (const __nv_bfloat16*)z, scale, shift,
                                                                 (const __nv_bfloat16*)zr, scale_r, shift_r, C, T, pitch,
                                                                 lens, relu, (__nv_bfloat16*)y, rows, (const float*)nullptr));
  TS_LAUNCH_CHECK("bn_apply_kernel");
  return TS_OK;
}

extern "C" int ts_bn_bwd_reduce(const void* dy, const void* y, const void* z, const void* zr, int B, int C, int T,
                                int pitch, int relu, float* sums, const float* mask_scale, const float* mask_shift,
                                const float* mask_scale_r, const float* mask_shift_r, void* stream) {
  TS_REQUIRE(dy && z && sums, TS_ERR_INVALID, "ts_bn_bwd_reduce: null pointer");
  TS_REQUIRE(!relu || y || (mask_scale && mask_shift && (!zr || (mask_scale_r && mask_shift_r))), TS_ERR_INVALID,
             "ts_bn_bwd_reduce: relu needs y or the forward scale/shift to rebuild the mask");
  TS_REQUIRE(B > 0 && C > 0 && T > 0 && pitch >= T && pitch % 8 == 0, TS_ERR_INVALID, "ts_bn_bwd_reduce: bad sizes");
  const long long rows = (long long)B * C;
  TS_CUDA(launch_pdl(train::bn_bwd_reduce_kernel, dim3((unsigned)ceil_div64(rows, train::RW)), dim3(train::RW * 32), 0, (cudaStream_t)stream, (option_pdl() & 2) != 0,
                     
      (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (const __nv_bfloat16*)z, (const __nv_bfloat16*)zr, T, pitch, rows,
      relu, sums, C, mask_scale, mask_shift, mask_scale_r, mask_shift_r, (const float*)nullptr));
  TS_LAUNCH_CHECK("bn_bwd_reduce_kernel");
  return TS_OK;
}

extern "C" int ts_bn_bwd_apply(const void* dy, const void* y, const void* z, const void* zr, const float* coef,
                               const float* coef_r, int B, int C, int T, int pitch, int relu, void* dz, void* dzr,
                               const float* mask_scale, const float* mask_shift, const float* mask_scale_r,
                               const float* mask_shift_r, void* stream) {
  TS_REQUIRE(dy && z && coef && dz, TS_ERR_INVALID, "ts_bn_bwd_apply: null pointer");
  TS_REQUIRE(!relu || y || (mask_scale && mask_shift && (!zr || (mask_scale_r && mask_shift_r))), TS_ERR_INVALID,
             "ts_bn_bwd_apply: relu needs y or the forward scale/shift to rebuild the mask");
  TS_REQUIRE((zr == nullptr) == (coef_r == nullptr) && (zr == nullptr) == (dzr == nullptr), TS_ERR_INVALID,
             "ts_bn_bwd_apply: residual operands disagree");
  TS_REQUIRE(B > 0 && C > 0 && T > 0 && pitch >= T && pitch % 8 == 0, TS_ERR_INVALID, "ts_bn_bwd_apply: bad sizes");
  const long long rows = (long long)B * C;
  TS_REQUIRE(rows < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_bn_bwd_apply: too many rows");
  dim3 grid((unsigned)rows, ceil_div(pitch / 8, 128));
  TS_CUDA(launch_pdl(train::bn_bwd_apply_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, (option_pdl() & 2) != 0,
                     
      (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (const __nv_bfloat16*)z, (const __nv_bfloat16*)zr, coef, coef_r, C,
      T, pitch, relu, (__nv_bfloat16*)dz, (__nv_bfloat16*)dzr, rows, mask_scale, mask_shift, mask_scale_r, mask_shift_r,
      (const float*)nullptr, (const float*)nullptr));
  TS_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return TS_OK;
}

namespace ts {
namespace train {

// One warp per channel: lanes stride over the NB partials (independent loads), shuffle-reduce in double.
__global__ void bn_finalize_kernel(const float* __restrict__ part, int NB, int slots, int C, double n,
                                   const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* __restrict__ rmean,
                                   float* __restrict__ rvar, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ inv_out) {
  pdl_launch_dependents();   // PDL: the next kernel may start its prologue; then wait for the previous grid
  pdl_wait();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int i = lane; i < NB * slots; i += 32) {   // part[b, c, slot, 2]
    const int b = i / slots, sl = i - b * slots;
    const float2 v = *reinterpret_cast<const float2*>(part + (((size_t)b * C + c) * slots + sl) * 2);
    s0 += v.x;
    s1 += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if (lane != 0) return;
  const double mean = s0 / n;
  double var = s1 / n - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const double inv = 1.0 / sqrt(var + (double)eps);
  const double sc = (double)gamma[c] * inv;
  scale[c] = (float)sc;
  shift[c] = (float)((double)beta[c] - mean * sc);
  mean_out[c] = (float)mean;
  inv_out[c] = (float)inv;
  if (rmean != nullptr) {
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)(var * n / (n > 1.0 ? n - 1.0 : 1.0));
  }
}

__global__ void bn_bwd_coef_kernel(const float* __restrict__ part, int NB, int C, int which, double n,
                                   const float* __restrict__ gamma, const float* __restrict__ mean,
                                   const float* __restrict__ inv, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                   float* __restrict__ coef) {
  pdl_launch_dependents();   // PDL: the next kernel may start its prologue; then wait for the previous grid
  pdl_wait();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int b = lane; b < NB; b += 32) {
    const float* p = part + ((size_t)b * C + c) * 3;
    s0 += p[0];
    s1 += p[which];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if (lane != 0) return;
  const double m = mean[c], iv = inv[c], g = gamma[c];
  const double dg = iv * (s1 - m * s0);
  const double a = g * iv;
  const double b2 = -g * iv * iv * dg / n;
  const double c0 = -a * s0 / n - b2 * m;
  dgamma[c] = (float)dg;
  dbeta[c] = (float)s0;
  coef[c * 3 + 0] = (float)a;
  coef[c * 3 + 1] = (float)b2;
  coef[c * 3 + 2] = (float)c0;
}

}  // namespace train
}  // namespace ts

extern "C" int ts_bn_finalize(const float* part, int NB, int slots, int C, double n, const float* gamma, const float* beta, float eps,
                              float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                              float* mean, float* inv, void* stream) {
  TS_REQUIRE(part && gamma && beta && scale && shift && mean && inv, TS_ERR_INVALID, "ts_bn_finalize: null pointer");
  TS_REQUIRE(NB > 0 && slots > 0 && C > 0 && n >= 1.0 && (running_mean == nullptr) == (running_var == nullptr), TS_ERR_INVALID,
             "ts_bn_finalize: bad sizes");
  TS_CUDA(launch_pdl(train::bn_finalize_kernel, dim3(ceil_div(C, 4)), dim3(128), 0, (cudaStream_t)stream, (option_pdl() & 2) != 0,
                     part, NB, slots, C, n, gamma, beta, eps, momentum,
                                                                               running_mean, running_var, scale, shift,
                                                                               mean, inv));
  TS_LAUNCH_CHECK("bn_finalize_kernel");
  return TS_OK;
}

extern "C" int ts_bn_bwd_coef(const float* part, int NB, int C, int which, double n, const float* gamma, const float* mean,
                              const float* inv, float* dgamma, float* dbeta, float* coef, void* stream) {
  TS_REQUIRE(part && gamma && mean && inv && dgamma && dbeta && coef, TS_ERR_INVALID, "ts_bn_bwd_coef: null pointer");
  TS_REQUIRE(NB > 0 && C > 0 && n >= 1.0 && (which == 1 || which == 2), TS_ERR_INVALID, "ts_bn_bwd_coef: bad arguments");
  TS_CUDA(launch_pdl(train::bn_bwd_coef_kernel, dim3(ceil_div(C, 4)), dim3(128), 0, (cudaStream_t)stream, (option_pdl() & 2) != 0,
                     part, NB, C, which, n, gamma, mean, inv,
                                                                               dgamma, dbeta, coef));
  TS_LAUNCH_CHECK("bn_bwd_coef_kernel");
  return TS_OK;
}

namespace ts {
namespace train {

// One launch prepares every weight operand of a training step from the fp32 master weights.  Table row (8 x int64):
//   src, dst, dstT, rows, cols, ldT, kind, first_tile
// kind 0 (pointwise / decoder weight [rows, cols] f32): dst = bf16 copy [rows, cols], dstT = bf16 transpose [cols, ldT]
// kind 1 (depthwise taps [rows, cols] f32): dst = taps rounded to bf16 kept as f32, dstT = the same, flipped along cols
// Tiles are 32 x 32; a CTA finds its table row by binary search over first_tile.
__global__ void __launch_bounds__(256)
prep_weights_kernel(const long long* __restrict__ tab, int n) {
  pdl_launch_dependents();   // PDL: the next kernel may start its prologue; then wait for the previous grid
  pdl_wait();
  __shared__ float tile[32][33];
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid * 8 + 7] <= (long long)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const long long* e = tab + lo * 8;
  const float* src = reinterpret_cast<const float*>(e[0]);
  const int rows = (int)e[3], cols = (int)e[4], ldT = (int)e[5], kind = (int)e[6];
  const int t = (int)((long long)blockIdx.x - e[7]);
  const int tc = (cols + 31) >> 5;
  const int r0 = (t / tc) << 5, c0 = (t % tc) << 5;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (kind == 0) {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e[1]);
    __nv_bfloat16* dstT = reinterpret_cast<__nv_bfloat16*>(e[2]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + 8 * i, c = c0 + tx;
      float v = 0.f;
      if (r < rows && c < cols) {
        v = src[(size_t)r * cols + c];
        dst[(size_t)r * cols + c] = __float2bfloat16(v);
      }
      tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + ty + 8 * i, r = r0 + tx;
      if (r < rows && c < cols) dstT[(size_t)c * ldT + r] = __float2bfloat16(tile[tx][ty + 8 * i]);
    }
  } else {
    float* dst = reinterpret_cast<float*>(e[1]);
    float* dstF = reinterpret_cast<float*>(e[2]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + 8 * i, c = c0 + tx;
      if (r < rows && c < cols) {
        const float v = __bfloat162float(__float2bfloat16(src[(size_t)r * cols + c]));
        dst[(size_t)r * cols + c] = v;
        dstF[(size_t)r * cols + (cols - 1 - c)] = v;
      }
    }
  }
}

}  // namespace train
}  // namespace ts

namespace ts {
namespace train {

// block-wide sum of two doubles (256 threads); result valid in every thread
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sm /*[16]*/) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) {
    sm[w] = a;
    sm[8 + w] = b;
  }
  __syncthreads();
  a = 0.0;
  b = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a += sm[i];
    b += sm[8 + i];
  }
}

struct BnSide {          // one BatchNorm branch of the fused forward kernel
  const __nv_bfloat16* z;
  const float* part;     // [NB, C, slots, 2]
  int NB, slots;
  const float* gamma;
  const float* beta;
  float* rmean;          // nullable: running statistics updated in place
  float* rvar;
  float* out;            // [4, C]: scale, shift, mean, inv (consumed by the backward pass)
  float eps, momentum;
};

// Fused BatchNorm finalize + apply: one CTA per (channel, 2048-frame chunk).  The CTA reduces the partial sums of its
// channel once (double), derives scale / shift, and then streams all B rows of the channel:
//   y = act(z * scale + shift (+ zr * scale_r + shift_r)), frames >= lens[b] and the pad stored as zero.
// blockIdx.y == 0 also publishes (scale, shift, mean, inv) and updates the running statistics.
__global__ void __launch_bounds__(256)
bn_apply_fused_kernel(BnSide m, BnSide r, int has_res, int B, int C, int T, int pitch, double n,
                      const int32_t* __restrict__ lens, int relu, __nv_bfloat16* __restrict__ y) {
  __shared__ double red[16];
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x;
  float sc[2] = {0.f, 0.f}, sh[2] = {0.f, 0.f};
  for (int side = 0; side < 1 + has_res; ++side) {
    const BnSide& q = side == 0 ? m : r;
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < q.NB * q.slots; i += 256) {
      const int b = i / q.slots, sl = i - b * q.slots;
      const float2 v = *reinterpret_cast<const float2*>(q.part + (((size_t)b * C + c) * q.slots + sl) * 2);
      s0 += v.x;
      s1 += v.y;
    }
    block_sum2(s0, s1, red);
    const double mean = s0 / n;
    double var = s1 / n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double inv = 1.0 / sqrt(var + (double)q.eps);
    const double scd = (double)q.gamma[c] * inv;
    sc[side] = (float)scd;
    sh[side] = (float)((double)q.beta[c] - mean * scd);
    if (blockIdx.y == 0 && threadIdx.x == 0) {
      q.out[c] = sc[side];
      q.out[C + c] = sh[side];
      q.out[2 * C + c] = (float)mean;
      q.out[3 * C + c] = (float)inv;
      if (q.rmean != nullptr) {
        q.rmean[c] = (1.f - q.momentum) * q.rmean[c] + q.momentum * (float)mean;
        q.rvar[c] = (1.f - q.momentum) * q.rvar[c] + q.momentum * (float)(var * n / (n > 1.0 ? n - 1.0 : 1.0));
      }
    }
  }
  // stream the channel: the B x pitch/8 vectors are flattened so that all 256 threads stay busy for any pitch
  const int pv = pitch >> 3, total = B * pv;
  const int per = (total + gridDim.y - 1) / gridDim.y;
  const int i1 = min(total, (int)(blockIdx.y + 1) * per);
  for (int i = blockIdx.y * per + threadIdx.x; i < i1; i += 256) {
    const int b = i / pv, t = (i - b * pv) << 3;
    const size_t off = ((size_t)b * C + c) * pitch + t;
    const int lim = lens ? min(T, max(lens[b], 0)) : T;
    float v[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(m.z + off), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(v[j], sc[0], sh[0]);
    if (has_res) {
      unpack8(*reinterpret_cast<const uint4*>(r.z + off), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += fmaf(v[j], sc[1], sh[1]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (relu) o[j] = fmaxf(o[j], 0.f);
      if (t + j >= lim) o[j] = 0.f;
    }
    *reinterpret_cast<uint4*>(y + off) = pack8(o);
  }
}

struct BnBwdSide {
  const __nv_bfloat16* z;
  const float* gamma;
  const float* stats;    // [4, C] from the forward pass: scale, shift, mean, inv
  float* dgamma;
  float* dbeta;
  __nv_bfloat16* dz;
};

// Fused BatchNorm backward coefficients + apply: one CTA per (channel, 2048-frame chunk).  From the partial sums of
// ts_bn_bwd_reduce (sums [B, C, 3]) the CTA derives dgamma, dbeta and the coefficients of dz = a dym + b z + c0 for its
// channel, then streams all B rows.  The ReLU mask is rebuilt from z (and zr) with the forward scale / shift.
__global__ void __launch_bounds__(256)
bn_bwd_apply_fused_kernel(const __nv_bfloat16* __restrict__ dy, BnBwdSide m, BnBwdSide r, int has_res,
                          const float* __restrict__ sums, int B, int C, int T, int pitch, double n, int relu) {
  __shared__ double red[16];
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, dummy = 0.0;
  for (int b = threadIdx.x; b < B; b += 256) {
    const float* p = sums + ((size_t)b * C + c) * 3;
    s0 += p[0];
    s1 += p[1];
    s2 += p[2];
  }
  block_sum2(s0, s1, red);
  if (has_res) block_sum2(s2, dummy, red);
  float ca[2], cb[2], cc[2], sc[2], sh[2];
  for (int side = 0; side < 1 + has_res; ++side) {
    const BnBwdSide& q = side == 0 ? m : r;
    const double mean = q.stats[2 * C + c], iv = q.stats[3 * C + c], g = q.gamma[c];
    const double sx = side == 0 ? s1 : s2;
    const double dg = iv * (sx - mean * s0);
    const double a = g * iv;
    const double b2 = -g * iv * iv * dg / n;
    ca[side] = (float)a;
    cb[side] = (float)b2;
    cc[side] = (float)(-a * s0 / n - b2 * mean);
    sc[side] = q.stats[c];
    sh[side] = q.stats[C + c];
    if (blockIdx.y == 0 && threadIdx.x == 0) {
      q.dgamma[c] = (float)dg;
      q.dbeta[c] = (float)s0;
    }
  }
  const int pv = pitch >> 3, total = B * pv;
  const int per = (total + gridDim.y - 1) / gridDim.y;
  const int i1 = min(total, (int)(blockIdx.y + 1) * per);
  for (int i = blockIdx.y * per + threadIdx.x; i < i1; i += 256) {
    const int b = i / pv, t = (i - b * pv) << 3;
    const size_t off = ((size_t)b * C + c) * pitch + t;
    float g[8], zz[8], z2[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(dy + off), g);
    unpack8(*reinterpret_cast<const uint4*>(m.z + off), zz);
    if (has_res) unpack8(*reinterpret_cast<const uint4*>(r.z + off), z2);
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float pre = fmaf(zz[j], sc[0], sh[0]);
        if (has_res) pre += fmaf(z2[j], sc[1], sh[1]);
        if (!(pre > 0.f)) g[j] = 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (t + j < T) ? fmaf(ca[0], g[j], fmaf(cb[0], zz[j], cc[0])) : 0.f;
    *reinterpret_cast<uint4*>(m.dz + off) = pack8(o);
    if (has_res) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (t + j < T) ? fmaf(ca[1], g[j], fmaf(cb[1], z2[j], cc[1])) : 0.f;
      *reinterpret_cast<uint4*>(r.dz + off) = pack8(o);
    }
  }
}

}  // namespace train
}  // namespace ts

extern "C" int ts_bn_apply_fused(const void* z, const float* part, int NB, int slots, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, float* stats_out, const void* zr,
                                 const float* part_r, int NB_r, int slots_r, const float* gamma_r, const float* beta_r,
                                 float* running_mean_r, float* running_var_r, float* stats_out_r, float eps,
                                 float momentum, int B, int C, int T, int pitch, const int32_t* lens, int relu, void* y,
                                 void* stream) {
  TS_REQUIRE(z && part && gamma && beta && stats_out && y, TS_ERR_INVALID, "ts_bn_apply_fused: null pointer");
  TS_REQUIRE(!zr || (part_r && gamma_r && beta_r && stats_out_r), TS_ERR_INVALID, "ts_bn_apply_fused: residual operands");
  TS_REQUIRE(B > 0 && C > 0 && C <= 65535 && T > 0 && pitch >= T && pitch % 8 == 0 && NB > 0 && slots > 0, TS_ERR_INVALID,
             "ts_bn_apply_fused: bad sizes");
  TS_REQUIRE((running_mean == nullptr) == (running_var == nullptr), TS_ERR_INVALID, "ts_bn_apply_fused: running stats");
  train::BnSide m{(const __nv_bfloat16*)z, part, NB, slots, gamma, beta, running_mean, running_var, stats_out, eps, momentum};
  train::BnSide r{(const __nv_bfloat16*)zr, part_r, NB_r, slots_r, gamma_r, beta_r, running_mean_r, running_var_r,
                  stats_out_r, eps, momentum};
  // grid.y: split a channel only when it is long (>= 16 vectors per thread and CTA)
  const int ny = std::max(1, std::min(64, (B * (pitch / 8)) / (256 * 16)));
  dim3 grid(C, ny);
  TS_CUDA(launch_pdl(train::bn_apply_fused_kernel, grid, dim3(256), 0, (cudaStream_t)stream, (option_pdl() & 2) != 0, m, r,
                     zr != nullptr ? 1 : 0, B, C, T, pitch, (double)B * (double)T, lens, relu, (__nv_bfloat16*)y));
  TS_LAUNCH_CHECK("bn_apply_fused_kernel");
  return TS_OK;
}

extern "C" int ts_bn_bwd_apply_fused(const void* dy, const void* z, const float* gamma, const float* stats, float* dgamma,
                                     float* dbeta, void* dz, const void* zr, const float* gamma_r, const float* stats_r,
                                     float* dgamma_r, float* dbeta_r, void* dzr, const float* sums, int B, int C, int T,
                                     int pitch, int relu, void* stream) {
  TS_REQUIRE(dy && z && gamma && stats && dgamma && dbeta && dz && sums, TS_ERR_INVALID, "ts_bn_bwd_apply_fused: null pointer");
  TS_REQUIRE(!zr || (gamma_r && stats_r && dgamma_r && dbeta_r && dzr), TS_ERR_INVALID, "ts_bn_bwd_apply_fused: residual operands");
  TS_REQUIRE(B > 0 && C > 0 && C <= 65535 && T > 0 && pitch >= T && pitch % 8 == 0, TS_ERR_INVALID,
             "ts_bn_bwd_apply_fused: bad sizes");
  train::BnBwdSide m{(const __nv_bfloat16*)z, gamma, stats, dgamma, dbeta, (__nv_bfloat16*)dz};
  train::BnBwdSide r{(const __nv_bfloat16*)zr, gamma_r, stats_r, dgamma_r, dbeta_r, (__nv_bfloat16*)dzr};
  const int ny = std::max(1, std::min(64, (B * (pitch / 8)) / (256 * 16)));
  dim3 grid(C, ny);
  TS_CUDA(launch_pdl(train::bn_bwd_apply_fused_kernel, grid, dim3(256), 0, (cudaStream_t)stream, (option_pdl() & 2) != 0,
                     (const __nv_bfloat16*)dy, m, r, zr != nullptr ? 1 : 0, sums, B, C, T, pitch, (double)B * (double)T, relu));
  TS_LAUNCH_CHECK("bn_bwd_apply_fused_kernel");
  return TS_OK;
}

extern "C" int ts_prep_weights(const long long* table, int n_entries, long long total_tiles, void* stream) {
  TS_REQUIRE(table != nullptr && n_entries > 0 && total_tiles > 0 && total_tiles < (1ll << 31), TS_ERR_INVALID,
             "ts_prep_weights: bad arguments");
  TS_CUDA(launch_pdl(train::prep_weights_kernel, dim3((unsigned)total_tiles), dim3(256), 0, (cudaStream_t)stream, (option_pdl() & 2) != 0,
                     table, n_entries));
  TS_LAUNCH_CHECK("prep_weights_kernel");
  return TS_OK;
}

int launch_dw_wgrad_tiled(const __nv_bfloat16* da, int T_out, int pitch_out, const __nv_bfloat16* x, int T_in, int pitch_in,
                          const int32_t* len_in, int B, int C, int K, int D, int P, int bchunk, float* part,
                          cudaStream_t st);

int launch_dw_wgrad_mma(const __nv_bfloat16* da, int pitch_out, const __nv_bfloat16* x, int pitch_in, int B, int C, int K,
                        int D, int P, float* out, cudaStream_t st);

extern "C" int ts_dw_wgrad(const void* da, int T_out, int pitch_out, const void* x, int T_in, int pitch_in,
                           const int32_t* len_in, int B, int C, int K, int S, int D, int P, int bchunk, int flags,
                           float* part, void* stream) {
  TS_REQUIRE(da && x && part, TS_ERR_INVALID, "ts_dw_wgrad: null pointer");
  TS_REQUIRE(B > 0 && C > 0 && K > 0 && K <= 256 && S > 0 && D > 0 && P >= 0 && bchunk > 0, TS_ERR_INVALID,
             "ts_dw_wgrad: bad sizes (K <= 256)");
  TS_REQUIRE(pitch_in % 8 == 0 && pitch_out % 8 == 0 && pitch_in >= T_in && pitch_out >= T_out, TS_ERR_INVALID,
             "ts_dw_wgrad: pitches must be multiples of 8 frames and >= T");
  if (S == 1 && T_in == T_out && bchunk >= B && (flags & TS_DW_INPUT_PREMASKED) && option_dw_mma()) {
    // tensor-core kernel (dwwgrad_mma.cu): whole batch per channel, writes part[0] = the final gradient
    const int rc = launch_dw_wgrad_mma((const __nv_bfloat16*)da, pitch_out, (const __nv_bfloat16*)x, pitch_in, B, C, K, D, P,
                                       part, (cudaStream_t)stream);
    if (rc != TS_ERR_UNSUPPORTED) return rc;
  }
  if (S == 1) {   // register-tiled kernel (dwwgrad.cu); the kernel below is the generic fallback (strided stem)
    const int rc = launch_dw_wgrad_tiled((const __nv_bfloat16*)da, T_out, pitch_out, (const __nv_bfloat16*)x, T_in, pitch_in,
                                         len_in, B, C, K, D, P, bchunk, part, (cudaStream_t)stream);
    if (rc != TS_ERR_UNSUPPORTED) return rc;
  }
  TS_REQUIRE((T_out - 1) * S + (K - 1) * D - P < pitch_in + K * D + 8, TS_ERR_INVALID, "ts_dw_wgrad: window exceeds row");
  const size_t smem = (size_t)(pitch_in + 2 * (K * D + 8) + pitch_out) * sizeof(float);
  TS_REQUIRE(smem <= 200 * 1024, TS_ERR_UNSUPPORTED, "ts_dw_wgrad: rows too long for shared memory");
  if (smem > 48 * 1024)
    TS_CUDA(cudaFuncSetAttribute(train::dw_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(C, ceil_div(B, bchunk));
  train::dw_wgrad_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)da, T_out, pitch_out,
                                                                    (const __nv_bfloat16*)x, T_in, pitch_in, len_in, B, C,
                                                                    K, S, D, P, bchunk, part);
  TS_LAUNCH_CHECK("dw_wgrad_kernel");
  return TS_OK;
}

// ---- SqueezeExcite variants (Citrinet training): the main branch of the block's last sub-block is scaled by gate[b, c] ----
extern "C" int ts_bn_apply_se(const void* z, const float* scale, const float* shift, const void* zr, const float* scale_r,
                              const float* shift_r, const float* gate, int B, int C, int T, int pitch, const int32_t* lens,
                              int relu, void* y, void* stream) {
  TS_REQUIRE(z && scale && shift && gate && y, TS_ERR_INVALID, "ts_bn_apply_se: null pointer");
  TS_REQUIRE((zr == nullptr) == (scale_r == nullptr) && (zr == nullptr) == (shift_r == nullptr), TS_ERR_INVALID,
             "ts_bn_apply_se: residual operands disagree");
  TS_REQUIRE(B > 0 && C > 0 && T > 0 && pitch >= T && pitch % 8 == 0, TS_ERR_INVALID, "ts_bn_apply_se: bad sizes");
  const long long rows = (long long)B * C;
  TS_REQUIRE(rows < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_bn_apply_se: too many rows");
  dim3 grid((unsigned)rows, ceil_div(pitch / 8, 128));
  TS_CUDA(launch_pdl(train::bn_apply_kernel, grid, dim3(128), 0, (cudaStream_t)stream, (option_pdl() & 2) != 0,
                     (const __nv_bfloat16*)z, scale, shift, (const __nv_bfloat16*)zr, scale_r, shift_r, C, T, pitch, lens, relu,
                     (__nv_bfloat16*)y, rows, gate));
  TS_LAUNCH_CHECK("bn_apply_kernel");
  return TS_OK;
}

extern "C" int ts_bn_bwd_reduce_se(const void* dy, const void* z, const void* zr, int B, int C, int T, int pitch, int relu,
                                   float* sums, const float* mask_scale, const float* mask_shift,
                                   const float* mask_scale_r, const float* mask_shift_r, const float* gate, void* stream) {
  TS_REQUIRE(dy && z && sums && mask_scale && mask_shift && gate, TS_ERR_INVALID, "ts_bn_bwd_reduce_se: null pointer");
  TS_REQUIRE(!zr || (mask_scale_r && mask_shift_r), TS_ERR_INVALID, "ts_bn_bwd_reduce_se: residual operands");
  TS_REQUIRE(B > 0 && C > 0 && T > 0 && pitch >= T && pitch % 8 == 0, TS_ERR_INVALID, "ts_bn_bwd_reduce_se: bad sizes");
  const long long rows = (long long)B * C;
  TS_CUDA(launch_pdl(train::bn_bwd_reduce_kernel, dim3((unsigned)ceil_div64(rows, train::RW)), dim3(train::RW * 32), 0,
                     (cudaStream_t)stream, (option_pdl() & 2) != 0, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)nullptr,
                     (const __nv_bfloat16*)z, (const __nv_bfloat16*)zr, T, pitch, rows, relu, sums, C, mask_scale, mask_shift,
                     mask_scale_r, mask_shift_r, gate));
  TS_LAUNCH_CHECK("bn_bwd_reduce_kernel");
  return TS_OK;
}

extern "C" int ts_bn_bwd_apply_se(const void* dy, const void* z, const void* zr, const float* coef, const float* coef_r, int B,
                                  int C, int T, int pitch, int relu, void* dz, void* dzr, const float* mask_scale,
                                  const float* mask_shift, const float* mask_scale_r, const float* mask_shift_r,
                                  const float* gate, const float* addc, void* stream) {
  TS_REQUIRE(dy && z && coef && dz && mask_scale && mask_shift && gate && addc, TS_ERR_INVALID,
             "ts_bn_bwd_apply_se: null pointer");
  TS_REQUIRE((zr == nullptr) == (coef_r == nullptr) && (zr == nullptr) == (dzr == nullptr), TS_ERR_INVALID,
             "ts_bn_bwd_apply_se: residual operands disagree");
  TS_REQUIRE(!zr || (mask_scale_r && mask_shift_r), TS_ERR_INVALID, "ts_bn_bwd_apply_se: residual mask operands");
  TS_REQUIRE(B > 0 && C > 0 && T > 0 && pitch >= T && pitch % 8 == 0, TS_ERR_INVALID, "ts_bn_bwd_apply_se: bad sizes");
  const long long rows = (long long)B * C;
  TS_REQUIRE(rows < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_bn_bwd_apply_se: too many rows");
  dim3 grid((unsigned)rows, ceil_div(pitch / 8, 128));
  TS_CUDA(launch_pdl(train::bn_bwd_apply_kernel, grid, dim3(128), 0, (cudaStream_t)stream, (option_pdl() & 2) != 0,
                     (const __nv_bfloat16*)dy, (const __nv_bfloat16*)nullptr, (const __nv_bfloat16*)z,
                     (const __nv_bfloat16*)zr, coef, coef_r, C, T, pitch, relu, (__nv_bfloat16*)dz, (__nv_bfloat16*)dzr, rows,
                     mask_scale, mask_shift, mask_scale_r, mask_shift_r, gate, addc));
  TS_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return TS_OK;
}

// ---- AdamW over every parameter of the model in ONE launch ------------------------------------------------------------
// torch.optim.AdamW's update (decoupled weight decay, bias-corrected moments):
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// Gradients and both moments live in flat fp32 buffers (same element order); the parameters stay where the module keeps
// them, addressed through a device table {param pointer, offset in the flat buffers, elements, first tile}; tiles are 1024
// elements, a CTA finds its row by binary search (like prep_weights_kernel).
namespace ts {
namespace train {
__global__ void __launch_bounds__(256)
adamw_kernel(const long long* __restrict__ tab, int n, const float* __restrict__ grad, float* __restrict__ m,
             float* __restrict__ v, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  pdl_launch_dependents();
  pdl_wait();
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid * 4 + 3] <= (long long)blockIdx.x) lo = mid;
    else hi = mid - 1;
  }
  const long long* e = tab + lo * 4;
  float* p = reinterpret_cast<float*>(e[0]);
  const long long off = e[1], cnt = e[2];
  const long long i0 = ((long long)blockIdx.x - e[3]) * 1024 + threadIdx.x * 4;
  const float decay = 1.f - lr * wd, step = lr / bc1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long i = i0 + j;
    if (i < cnt) {
      const float g = grad[off + i];
      const float mi = b1 * m[off + i] + (1.f - b1) * g;
      const float vi = b2 * v[off + i] + (1.f - b2) * g * g;
      m[off + i] = mi;
      v[off + i] = vi;
      p[i] = p[i] * decay - step * mi / (sqrtf(vi) / bc2_sqrt + eps);
    }
  }
}
}  // namespace train
}  // namespace ts

extern "C" int ts_adamw(const long long* table, int n_entries, long long total_tiles, const float* grad, float* exp_avg,
                        float* exp_avg_sq, float lr, float beta1, float beta2, float eps, float weight_decay,
                        float bias_correction1, float bias_correction2, void* stream) {
  TS_REQUIRE(table && grad && exp_avg && exp_avg_sq, TS_ERR_INVALID, "ts_adamw: null pointer");
  TS_REQUIRE(n_entries > 0 && total_tiles > 0 && total_tiles < (1ll << 31), TS_ERR_INVALID, "ts_adamw: bad sizes");
  TS_REQUIRE(bias_correction1 > 0.f && bias_correction2 > 0.f, TS_ERR_INVALID, "ts_adamw: bias corrections must be positive");
  TS_CUDA(launch_pdl(train::adamw_kernel, dim3((unsigned)total_tiles), dim3(256), 0, (cudaStream_t)stream,
                     (option_pdl() & 2) != 0, table, n_entries, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay,
                     bias_correction1, sqrtf(bias_correction2)));
  TS_LAUNCH_CHECK("adamw_kernel");
  return TS_OK;
}
