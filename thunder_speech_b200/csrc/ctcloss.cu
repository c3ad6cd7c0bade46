// CTC loss and its gradient w.r.t. the logits, log-softmax fused (calculate_ctc, src/thunder/ctc_loss.py:15-47:
// log_softmax over the vocabulary -> F.ctc_loss(reduction="mean", zero_infinity=True)).
//
//   logits  fp32 rows [B, V, pitch] (time contiguous: the decoder GEMM's output layout)
//   targets int64 [B, Lmax], target lengths int64 [B], input lengths int32 [B] (encoder output lengths)
//
// ctc_alpha_beta_kernel: one CTA per utterance.  lse[t] = logsumexp_v logits[v, t]; then 256 threads run the alpha
// recursion forwards in time while the other 256 run the beta recursion backwards (one __syncthreads per frame, states in
// double-buffered shared memory, the next frame's log-probability prefetched from global memory).  alpha / beta go to a
// scratch [B, T, Sp] for the gradient kernel; nll[b] = -logaddexp(alpha_{len-1}(S-1), alpha_{len-1}(S-2)).
//
// ctc_grad_kernel: a CTA per (64 frames, utterance, 128 classes).  With y_t = softmax and the identity
//   sum_s alpha_t(s) beta_t(s) / y_t(l'_s) = P(l | x)   (both recursions include y_t(l'_s)),
//   d nll / d logit[c, t] = y_t(c) - sum_{s: l'_s = c} exp(alpha_t(s) + beta_t(s) - lp_t(c) + nll)
// scaled by gscale / (B * max(target_len, 1)) ("mean" reduction) and zero for t >= len or a non-finite nll
// (zero_infinity).  The tile is staged in shared memory and written as bf16 rows [B, Vp, pitch] ready for the decoder's
// weight / input gradient GEMMs.
#include "ts_common.cuh"

namespace ts {
namespace ctc {

constexpr float NEG_INF = -INFINITY;
constexpr int ROLE_THREADS = 256;
constexpr int MAX_NS = 4;   // states per thread: S = 2 L + 1 <= 1024

__device__ __forceinline__ int clamp_ll(long long v, int lo, int hi) {
  return v < lo ? lo : (v > hi ? hi : (int)v);
}
__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == NEG_INF) return NEG_INF;
  return m + log1pf(expf(-fabsf(a - b)));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, b), c);
  if (m == NEG_INF) return NEG_INF;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));   // sum in [1, 3]: absolute error of __logf ~2^-21
}

__global__ void __launch_bounds__(2 * ROLE_THREADS)
ctc_alpha_beta_kernel(const float* __restrict__ logits, int V, int T, int pitch, const int32_t* __restrict__ in_len,
                      const int64_t* __restrict__ targets, int Lmax, const int64_t* __restrict__ tgt_len, int blank,
                      float* __restrict__ alpha, float* __restrict__ beta, int Sp, float* __restrict__ lse_out,
                      float* __restrict__ nll_out) {
  extern __shared__ float sm[];
  float* lse_s = sm;                        // [T]
  float* st = lse_s + T;                    // [2 roles][2 buffers][Sp + 4], index 2 + s (two -inf guards each side)
  int* lab = reinterpret_cast<int*>(st + 4 * (Sp + 4));   // [Sp]
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* lg = logits + (size_t)b * V * pitch;
  const int len = min(max(in_len[b], 0), T);
  const int L = clamp_ll(tgt_len[b], 0, Lmax);
  const int S = 2 * L + 1;

  for (int t = tid; t < len; t += blockDim.x) {
    float m = NEG_INF;
    for (int v = 0; v < V; ++v) m = fmaxf(m, lg[(size_t)v * pitch + t]);
    float s = 0.f;
    for (int v = 0; v < V; ++v) s += expf(lg[(size_t)v * pitch + t] - m);
    const float l = m + logf(s);
    lse_s[t] = l;
    lse_out[(size_t)b * T + t] = l;
  }
  for (int s = tid; s < Sp; s += blockDim.x) {
    int l = blank;
    if ((s & 1) && s < S) {
      l = clamp_ll(targets[(size_t)b * Lmax + (s >> 1)], 0, V - 1);
    }
    lab[s] = l;
  }
  for (int i = tid; i < 4 * (Sp + 4); i += blockDim.x) st[i] = NEG_INF;
  __syncthreads();
  if (len == 0) {
    if (tid == 0) nll_out[b] = INFINITY;
    return;
  }

  const int role = tid / ROLE_THREADS, lane = tid % ROLE_THREADS;   // 0: alpha (forwards), 1: beta (backwards)
  float* buf = st + role * 2 * (Sp + 4);
  float* out = (role == 0 ? alpha : beta) + (size_t)b * T * Sp;
  int mylab[MAX_NS];
  bool skip[MAX_NS];
  float nxt[MAX_NS];
#pragma unroll
  for (int i = 0; i < MAX_NS; ++i) {
    const int s = lane + i * ROLE_THREADS;
    mylab[i] = (s < S) ? lab[s] : blank;
    if (role == 0) skip[i] = (s < S) && (s & 1) && s >= 2 && lab[s - 2] != lab[s];
    else skip[i] = (s & 1) && (s + 2 < S) && lab[s + 2] != lab[s];
  }
  const int t_first = role == 0 ? 0 : len - 1, dt = role == 0 ? 1 : -1;
#pragma unroll
  for (int i = 0; i < MAX_NS; ++i) {
    const int s = lane + i * ROLE_THREADS;
    nxt[i] = (s < S) ? lg[(size_t)mylab[i] * pitch + t_first] : 0.f;
  }
  for (int step = 0; step < len; ++step) {
    const int t = t_first + step * dt;
    float cur_lp[MAX_NS];
#pragma unroll
    for (int i = 0; i < MAX_NS; ++i) cur_lp[i] = nxt[i] - lse_s[t];
    if (step + 1 < len) {
#pragma unroll
      for (int i = 0; i < MAX_NS; ++i) {
        const int s = lane + i * ROLE_THREADS;
        if (s < S) nxt[i] = lg[(size_t)mylab[i] * pitch + t + dt];
      }
    }
    const float* prev = buf + ((step + 1) & 1) * (Sp + 4) + 2;
    float* cur = buf + (step & 1) * (Sp + 4) + 2;
#pragma unroll
    for (int i = 0; i < MAX_NS; ++i) {
      const int s = lane + i * ROLE_THREADS;
      if (s < S) {
        float v;
        if (step == 0) {
          const bool init = role == 0 ? (s <= 1) : (s >= S - 2);
          v = init ? cur_lp[i] : NEG_INF;
        } else if (role == 0) {
          const float c = skip[i] ? prev[s - 2] : NEG_INF;
          const float a = lse3(prev[s], prev[s - 1], c);
          v = a == NEG_INF ? NEG_INF : a + cur_lp[i];
        } else {
          const float c = skip[i] ? prev[s + 2] : NEG_INF;
          const float a = lse3(prev[s], (s + 1 < S) ? prev[s + 1] : NEG_INF, c);
          v = a == NEG_INF ? NEG_INF : a + cur_lp[i];
        }
        cur[s] = v;
        out[(size_t)t * Sp + s] = v;
      }
    }
    // the two recursions are independent: each role synchronises on its own named barrier (ids 1 and 2)
    asm volatile("bar.sync %0, %1;" ::"r"(role + 1), "r"(ROLE_THREADS) : "memory");
  }
  if (tid == 0) {
    const float* last = st + ((len - 1) & 1) * (Sp + 4) + 2;   // alpha buffer of the final step
    const float ll = lse2(last[S - 1], S >= 2 ? last[S - 2] : NEG_INF);
    nll_out[b] = -ll;
  }
}

constexpr int GT = 64;    // frames per CTA of the gradient kernel
constexpr int CH = 128;   // classes per CTA (grid.z walks the vocabulary in chunks: any V)

// grid (ceil(pitch / GT), B, ceil(V / CH)).  Phase A: tile[c][t] = softmax_t(c) * scale for the chunk's classes.  Phase B:
// one thread per frame walks the S states in order and subtracts the occupancy exp(alpha + beta - lp + nll) * scale from the
// row of the state's label -- sequential per frame, so the sum over repeated labels has a fixed order (deterministic) and
// the cost is O(S) per frame instead of O(V * S).
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const float* __restrict__ logits, int V, int Vp, int T, int lpitch, int pitch,
                const int32_t* __restrict__ in_len,
                const int64_t* __restrict__ targets, int Lmax, const int64_t* __restrict__ tgt_len, int blank,
                const float* __restrict__ alpha, const float* __restrict__ beta, int Sp, const float* __restrict__ lse,
                const float* __restrict__ nll, float gscale, int B, float* __restrict__ loss_out,
                __nv_bfloat16* __restrict__ grad) {
  extern __shared__ float sm[];
  float* tile = sm;                                          // [CH][GT + 1]
  int* lab = reinterpret_cast<int*>(tile + CH * (GT + 1));   // [Sp]
  const int b = blockIdx.y, t0 = blockIdx.x * GT, c0 = blockIdx.z * CH, tid = threadIdx.x;
  const int nc = min(CH, V - c0);
  const int len = min(max(in_len[b], 0), T);
  const int L = clamp_ll(tgt_len[b], 0, Lmax);
  const int S = 2 * L + 1;
  const float n = nll[b];
  const bool ok = isfinite(n) && len > 0;
  const float scale = ok ? gscale / ((float)B * (float)max(L, 1)) : 0.f;
  if (blockIdx.x == 0 && blockIdx.z == 0 && tid == 0) loss_out[b] = ok ? n / (float)max(L, 1) : 0.f;
  for (int s = tid; s < Sp; s += blockDim.x) {
    int l = blank;
    if ((s & 1) && s < S) {
      l = clamp_ll(targets[(size_t)b * Lmax + (s >> 1)], 0, V - 1);
    }
    lab[s] = l;
  }
  const float* lg = logits + (size_t)b * V * lpitch;
  // phase A: softmax term (frames are the fast index: coalesced reads of the logit rows)
  for (int i = tid; i < nc * GT; i += blockDim.x) {
    const int c = i / GT, tt = i - c * GT, t = t0 + tt;
    float g = 0.f;
    if (ok && t < len) g = expf(lg[(size_t)(c0 + c) * lpitch + t] - lse[(size_t)b * T + t]) * scale;
    tile[c * (GT + 1) + tt] = g;
  }
  __syncthreads();
  // phase B: occupancy term, one thread per frame
  if (tid < GT) {
    const int t = t0 + tid;
    if (ok && t < len) {
      const float* a = alpha + ((size_t)b * T + t) * Sp;
      const float* be = beta + ((size_t)b * T + t) * Sp;
      const float ls = lse[(size_t)b * T + t];
      for (int s = 0; s < S; ++s) {
        const int c = lab[s] - c0;
        if (c >= 0 && c < nc) {
          const float lp = lg[(size_t)(c0 + c) * lpitch + t] - ls;
          tile[c * (GT + 1) + tid] -= expf(a[s] + be[s] - lp + n) * scale;
        }
      }
    }
  }
  __syncthreads();
  // rows [V..Vp) stay as the caller zero-initialised them; frames >= T up to the pitch are zeroed here
  const int tmax = min(GT, pitch - t0);
  for (int i = tid; i < nc * GT; i += blockDim.x) {
    const int c = i / GT, tt = i - c * GT;
    if (tt < tmax)
      grad[((size_t)b * Vp + c0 + c) * pitch + t0 + tt] = __float2bfloat16((t0 + tt < T) ? tile[c * (GT + 1) + tt] : 0.f);
  }
}

}  // namespace ctc
}  // namespace ts

using namespace ts;

extern "C" int ts_ctc_loss(const float* logits, int B, int V, int T, int pitch, const int32_t* in_len,
                           const int64_t* targets, int Lmax, const int64_t* tgt_len, int blank, float gscale, float* scratch,
                           long long scratch_floats, float* loss, void* grad, int Vp, int grad_pitch, void* stream) {
  TS_REQUIRE(logits && in_len && targets && tgt_len && scratch && loss && grad, TS_ERR_INVALID, "ts_ctc_loss: null pointer");
  TS_REQUIRE(B > 0 && V > 0 && T > 0 && pitch >= T && grad_pitch >= T && grad_pitch % 8 == 0 && Lmax > 0 && Vp >= V,
             TS_ERR_INVALID, "ts_ctc_loss: bad sizes");
  TS_REQUIRE(blank >= 0 && blank < V, TS_ERR_INVALID, "ts_ctc_loss: blank index outside the vocabulary");
  const int S = 2 * Lmax + 1;
  TS_REQUIRE(S <= ctc::MAX_NS * ctc::ROLE_THREADS, TS_ERR_UNSUPPORTED, "ts_ctc_loss: target length > 511");
  const int Sp = round_up(S, 4);
  const long long need = 2ll * B * T * Sp + (long long)B * T + B;
  TS_REQUIRE(scratch_floats >= need, TS_ERR_INVALID, "ts_ctc_loss: scratch too small (need 2*B*T*Sp + B*T + B floats)");
  float* alpha = scratch;
  float* beta = alpha + (size_t)B * T * Sp;
  float* lse = beta + (size_t)B * T * Sp;
  float* nll = lse + (size_t)B * T;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem1 = (size_t)(T + 4 * (Sp + 4) + Sp) * 4;
  TS_REQUIRE(smem1 <= 200 * 1024, TS_ERR_UNSUPPORTED, "ts_ctc_loss: sequence too long for shared memory");
  if (smem1 > 48 * 1024)
    TS_CUDA(cudaFuncSetAttribute(ctc::ctc_alpha_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  ctc::ctc_alpha_beta_kernel<<<B, 2 * ctc::ROLE_THREADS, smem1, st>>>(logits, V, T, pitch, in_len, targets, Lmax, tgt_len, blank,
                                                                      alpha, beta, Sp, lse, nll);
  TS_LAUNCH_CHECK("ctc_alpha_beta_kernel");
  const size_t smem2 = (size_t)(ctc::CH * (ctc::GT + 1) + Sp) * 4;
  if (smem2 > 48 * 1024)
    TS_CUDA(cudaFuncSetAttribute(ctc::ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  TS_REQUIRE(B <= 65535 && ceil_div(V, ctc::CH) <= 65535, TS_ERR_UNSUPPORTED, "ts_ctc_loss: batch / vocabulary too large");
  dim3 grid(ceil_div(grad_pitch, ctc::GT), B, ceil_div(V, ctc::CH));
  ctc::ctc_grad_kernel<<<grid, 256, smem2, st>>>(logits, V, Vp, T, pitch, grad_pitch, in_len, targets, Lmax, tgt_len, blank, alpha, beta, Sp,
                                                 lse, nll, gscale, B, loss, (__nv_bfloat16*)grad);
  TS_LAUNCH_CHECK("ctc_grad_kernel");
  return TS_OK;
}
