// Feature front-end kernels.
//
//   logmel_kernel          pre-emphasis + reflect padding + window + 512-point STFT + |.|^2 + sparse mel
//                          projection + log, one pass over the audio, nothing but the [B,nfilt,F] log-mel
//                          tensor is written.  Replaces PreEmphasisFilter/PowerSpectrum/MelScale
//                          (src/thunder/quartznet/transform.py:136-144,186-208,243-255).
//   normalize_rows_kernel  masked per-(batch, feature) normalisation, one warp per row held in registers.
//                          Replaces FeatureBatchNormalizer (transform.py:77-92, src/thunder/blocks.py:118-149).
//
// STFT: two real frames are packed into one complex 512-point FFT per warp.  Each lane owns 16 complex
// points; three radix-8 passes with two shared-memory transposes (layouts chosen bank-conflict free, see
// tools/fft_layout_proto.py which emulates this exact dataflow against numpy.fft).
#include "ts_common.cuh"
#include "fft8.cuh"

namespace ts {
namespace feat {

constexpr int NFFT = 512;
constexpr int NBINS = NFFT / 2 + 1;
constexpr int FT = 32;       // frames per CTA
constexpr int NWARPS = 8;    // 256 threads; each warp transforms FT / NWARPS = 4 frames (2 packed pairs)
constexpr int XBUF = 576;    // float2 elements per warp exchange buffer (max index 72*7+63 = 567)
constexpr int XS1 = 72;      // exchange 1: idx = 72*k2 + 8*n1 + n0
constexpr int XS0 = 66;      // exchange 2: idx = 66*n0 + 8*k1 + k2 (float2 elements: distinct mod 16 per half-warp)
constexpr int TILE_LD = FT + 1;
constexpr int MAX_NNZ = 2048;
constexpr int MAX_NFILT = 128;

struct SmemLayout {
  int raw, raw_stride, tw, mstart, mcount, moff, mw, xch, tile, total;   // offsets in floats
};

__host__ __device__ inline SmemLayout smem_layout(int hop, int nfilt, int nnz) {
  SmemLayout L;
  int o = 0;
  L.raw_stride = ((FT - 1) * hop + NFFT + 8 + 3) & ~3;   // span + previous sample + alignment slack
  L.raw = o;    o += 2 * L.raw_stride;                   // double buffered
  L.tw = o;     o += 2 * NFFT;                            // float2
  L.mstart = o; o += (nfilt + 3) & ~3;
  L.mcount = o; o += (nfilt + 3) & ~3;
  L.moff = o;   o += (nfilt + 3) & ~3;
  L.mw = o;     o += (nnz + 3) & ~3;
  L.xch = o;    o += NWARPS * 2 * XBUF;                   // float2 per warp
  L.tile = o;   o += nfilt * TILE_LD;
  L.total = o;
  return L;
}

__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- train-mode dither (DitherAudio, src/thunder/quartznet/transform.py:109-118: x + dither * N(0, 1)) ------------------
// Counter-based noise: the four standard normals of samples [4 g, 4 g + 4) of utterance b are Philox4x32-10(key = seed,
// counter = (g, b)) pushed through Box-Muller, so a sample gets the same noise from every frame / tile / boundary path
// that touches it and the result depends only on (seed, b, sample index) -- not on the launch geometry.  The stream is
// NOT torch's (the reference draws torch.randn_like; parity is statistical by construction).
__device__ __forceinline__ float4 dither_normal4(unsigned long long seed, uint32_t g, uint32_t b) {
  uint32_t c0 = g, c1 = b, c2 = 0x5eed5eedu, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0;
    c1 = l1;
    c2 = h0 ^ c3 ^ k1;
    c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  // uniforms in (0, 1): 24 random bits + 1/2 ulp
  const float u0 = ((float)(c0 >> 8) + 0.5f) * 5.9604644775390625e-08f, u1 = ((float)(c1 >> 8) + 0.5f) * 5.9604644775390625e-08f;
  const float u2 = ((float)(c2 >> 8) + 0.5f) * 5.9604644775390625e-08f, u3 = ((float)(c3 >> 8) + 0.5f) * 5.9604644775390625e-08f;
  const float r0 = sqrtf(-2.f * __logf(u0)), r1 = sqrtf(-2.f * __logf(u2));
  float s0, co0, s1, co1;
  __sincosf(6.283185307179586f * u1, &s0, &co0);
  __sincosf(6.283185307179586f * u3, &s1, &co1);
  return make_float4(r0 * co0, r0 * s0, r1 * co1, r1 * s1);
}
__device__ __forceinline__ float dither_normal(unsigned long long seed, int s, int b) {
  const float4 z = dither_normal4(seed, (uint32_t)s >> 2, (uint32_t)b);
  const int e = s & 3;
  return e == 0 ? z.x : e == 1 ? z.y : e == 2 ? z.z : z.w;
}

// ---- any other n_fft (the reference takes whatever torch.stft takes, transform.py:258-271) -----------------------------
// Direct DFT over the window support, one CTA per (utterance, GF frames): the windowed, pre-emphasised, reflect-padded
// samples of the frames and the twiddle table sit in shared memory, every thread owns bins k, k + 256, ... and walks the
// window once for all GF frames (the twiddle index advances by k modulo n_fft, no multiplication).  O(n_fft * win) per frame
// instead of O(n_fft log n_fft): a correct, self-contained fallback (n_fft = 1024, win 400: ~2 ms for 64 x 20 s), not a
// tuned kernel -- n_fft = 512, the only size the reference's models use, has the FFT kernel above.
template <int GF>
__global__ void __launch_bounds__(256)
logmel_generic_kernel(const float* __restrict__ audio, int N, int F, int n_fft, int hop, float preemph,
                      const float* __restrict__ window_full, int win_lo, int win_hi, const float2* __restrict__ twiddle,
                      const int32_t* __restrict__ mel_start, const int32_t* __restrict__ mel_count,
                      const int32_t* __restrict__ mel_off, const float* __restrict__ mel_w, int nfilt,
                      float* __restrict__ logmel, float dither, unsigned long long seed,
                      const unsigned long long* __restrict__ seed_dev) {
  extern __shared__ __align__(16) float smem[];
  const int wlen = win_hi - win_lo, nbins = n_fft / 2 + 1;
  float2* tw = reinterpret_cast<float2*>(smem);          // [n_fft]
  float* v = smem + 2 * n_fft;                            // [GF][wlen]
  float* pw = v + GF * wlen;                              // [GF][nbins]
  const int b = blockIdx.y, f0 = blockIdx.x * GF, tid = threadIdx.x;
  const float* x = audio + (size_t)b * N;
  if (dither != 0.f && seed_dev != nullptr) seed ^= *seed_dev;
  for (int i = tid; i < n_fft; i += 256) tw[i] = twiddle[i];
  for (int i = tid; i < GF * wlen; i += 256) {
    const int g = i / wlen, n = win_lo + (i - g * wlen);
    const int f = f0 + g;
    float val = 0.f;
    if (f < F) {
      const int s = f * hop - n_fft / 2 + n;
      const int r = s < 0 ? -s : (s >= N ? 2 * (N - 1) - s : s);      // reflect (torch.stft center=True)
      float xr = x[r], xp = r >= 1 ? x[r - 1] : 0.f;
      if (dither != 0.f) {
        xr = fmaf(dither, dither_normal(seed, r, b), xr);
        if (r >= 1) xp = fmaf(dither, dither_normal(seed, r - 1, b), xp);
      }
      val = window_full[n] * (r >= 1 ? xr - preemph * xp : xr);       // y[0] = x[0]
    }
    v[i] = val;
  }
  __syncthreads();
  for (int k = tid; k < nbins; k += 256) {
    float re[GF], im[GF];
#pragma unroll
    for (int g = 0; g < GF; ++g) re[g] = im[g] = 0.f;
    int idx = (int)(((long long)k * win_lo) % n_fft);
    for (int j = 0; j < wlen; ++j) {
      const float2 t = tw[idx];
#pragma unroll
      for (int g = 0; g < GF; ++g) {
        const float a = v[g * wlen + j];
        re[g] = fmaf(a, t.x, re[g]);
        im[g] = fmaf(a, t.y, im[g]);
      }
      idx += k;
      if (idx >= n_fft) idx -= n_fft;
    }
#pragma unroll
    for (int g = 0; g < GF; ++g) {
      const float m = sqrtf(re[g] * re[g] + im[g] * im[g]);             // sqrt then square, transform.py:205-207
      pw[g * nbins + k] = m * m;
    }
  }
  __syncthreads();
  for (int i = tid; i < GF * nfilt; i += 256) {
    const int m = i / GF, g = i - m * GF;                                // consecutive threads: consecutive frames
    if (f0 + g >= F) continue;
    const int st = mel_start[m], cnt = mel_count[m], off = mel_off[m];
    float a = 0.f;
    for (int j = 0; j < cnt; ++j) a = fmaf(mel_w[off + j], pw[g * nbins + st + j], a);
    logmel[((size_t)b * nfilt + m) * F + f0 + g] = logf(a + 5.9604644775390625e-08f);
  }
}

// Persistent kernel: each CTA loops over (utterance, 32-frame tile) work items.  The raw audio span of the NEXT item is
// fetched with cp.async into the other half of a double buffer while the current item is transformed.
//
// Buffer semantics: interior items hold RAW samples v[i] = x[a0 + i] and the pre-emphasis is folded into the window
// multiply ( w[n]*(x[n] - p*x[n-1]) = w[n]*v[n] - (p*w[n])*v[n-1] ); items that touch a row boundary (reflect padding,
// y[0] = x[0]) are filled synchronously with the already pre-emphasised, reflected samples and use p_eff = 0.
//
// kCentre320: window support is inside [96, 416) (win_length 320 centred in 512) => only FFT input slots q in [3, 12]
// (n = lane + 32 q) are non-zero for every lane, the rest are compile-time zeros.
template <bool kCentre320>
__global__ void __launch_bounds__(NWARPS * 32, 2)
logmel_kernel(const float* __restrict__ audio, int B, int N, int F, int hop, float preemph,
              const float* __restrict__ window_full, const float2* __restrict__ twiddle,
              const int32_t* __restrict__ mel_start, const int32_t* __restrict__ mel_count,
              const int32_t* __restrict__ mel_off, const float* __restrict__ mel_w, int nfilt, int nnz,
              float* __restrict__ logmel, int tiles_per_row, int num_items, float dither, unsigned long long seed,
              const unsigned long long* __restrict__ seed_dev) {
  extern __shared__ __align__(16) float smem[];
  if (dither != 0.f && seed_dev != nullptr) seed ^= *seed_dev;   // per-step state kept on the device (graph replays)
  const SmemLayout L = smem_layout(hop, nfilt, nnz);
  float2* tw = reinterpret_cast<float2*>(smem + L.tw);
  int* s_mstart = reinterpret_cast<int*>(smem + L.mstart);
  int* s_mcount = reinterpret_cast<int*>(smem + L.mcount);
  int* s_moff = reinterpret_cast<int*>(smem + L.moff);
  float* s_mw = smem + L.mw;
  float* tile = smem + L.tile;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  float2* xch = reinterpret_cast<float2*>(smem + L.xch) + warp * XBUF;

  // ---- per-CTA tables; per-lane window taps (and preemph * taps) in registers ----------------------------
  for (int i = tid; i < NFFT; i += NWARPS * 32) tw[i] = twiddle[i];
  for (int i = tid; i < nfilt; i += NWARPS * 32) {
    s_mstart[i] = mel_start[i];
    s_mcount[i] = mel_count[i];
    s_moff[i] = mel_off[i];
  }
  for (int i = tid; i < nnz; i += NWARPS * 32) s_mw[i] = mel_w[i];
  constexpr int Q0 = kCentre320 ? 3 : 0, Q1 = kCentre320 ? 13 : 16;
  float wq[Q1 - Q0], wpq[Q1 - Q0];
#pragma unroll
  for (int q = Q0; q < Q1; ++q) {
    wq[q - Q0] = window_full[lane + 32 * q];
    wpq[q - Q0] = preemph * wq[q - Q0];
  }
  // per-lane twiddles of pass 1: W64^{n1 k2}, n1 = lane/8 + 4h (k2 = 0 is 1)
  const int n0 = lane & 7;
  const int span = (FT - 1) * hop + NFFT;     // samples of one item (without the "previous sample")
  const bool vec_ok = (N % 4 == 0) && (hop % 4 == 0);

  // item -> (b, f0); fill buffer `buf` for item; returns p_eff selector through smem flag
  __shared__ int s_boundary[2];
  auto stage_item = [&](int item, int buf) {
    const int b = item / tiles_per_row, f0 = (item - b * tiles_per_row) * FT;
    const float* x = audio + (size_t)b * N;
    const int s0 = f0 * hop - NFFT / 2;       // signal index of frame f0's sample 0
    float* raw = smem + L.raw + buf * L.raw_stride;
    // layout: raw[4 + i] <-> signal index s0 + i  (i = -1 is the previous sample); 4 floats of slack keep 16 B alignment
    const bool interior = (s0 - 4 >= 0) && (s0 + span + 4 <= N);
    if (interior) {
      if (tid == 0) s_boundary[buf] = 0;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(raw);
      if (vec_ok) {  // s0 % 4 == 0: copy [s0 - 4, s0 + span + 4) as 16-byte chunks
        const float* src = x + s0 - 4;
        for (int i = tid; i < (span + 8) / 4; i += NWARPS * 32) cp_async_16(dst + 16 * i, src + 4 * i);
      } else {
        const float* src = x + s0 - 4;
        for (int i = tid + 3; i < span + 4; i += NWARPS * 32) cp_async_4(dst + 4 * i, src + i);
      }
    } else {
      if (tid == 0) s_boundary[buf] = 1;
      if (tid < 4) raw[tid] = 0.f;   // the "previous sample" slot is multiplied by p_eff = 0: keep it finite
      for (int i = tid; i < span; i += NWARPS * 32) {
        const int s = s0 + i;
        float v = 0.f;
        if (s > -NFFT / 2 - 1 && s < N + NFFT / 2) {
          const int r = s < 0 ? -s : (s >= N ? 2 * (N - 1) - s : s);  // reflect (torch.stft center=True)
          if (dither == 0.f) {
            v = (r >= 1) ? (x[r] - preemph * x[r - 1]) : x[0];          // y[0] = x[0]
          } else {   // the dithered signal is pre-emphasised: xd[r] = x[r] + dither * n(b, r)
            const float xr = fmaf(dither, dither_normal(seed, r, b), x[r]);
            v = (r >= 1) ? (xr - preemph * fmaf(dither, dither_normal(seed, r - 1, b), x[r - 1])) : xr;
          }
        }
        raw[4 + i] = v;
      }
    }
    cp_async_commit();
  };

  int item = blockIdx.x;
  if (item < num_items) stage_item(item, 0);
  for (int it = 0; item < num_items; item += gridDim.x, ++it) {
    const int buf = it & 1;
    const int next = item + gridDim.x;
    if (next < num_items) {
      stage_item(next, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // buffer `buf` (and, first time, the tables) visible; previous item's tile rows written out
    const int b = item / tiles_per_row, f0 = (item - b * tiles_per_row) * FT;
    const float* raw = smem + L.raw + buf * L.raw_stride + 4;
    const float pe = s_boundary[buf] ? 0.f : 1.f;
    if (dither != 0.f && !s_boundary[buf]) {   // train mode: add the noise to the staged raw samples (interior items; uniform)
      float* rw = smem + L.raw + buf * L.raw_stride;            // rw[i] <-> signal index s0 - 4 + i, i in [3, span + 4)
      const int sbase = f0 * hop - NFFT / 2 - 4;
      if ((sbase & 3) == 0) {                                   // groups of four aligned with the Philox counter
        for (int i4 = tid; i4 < (span + 8) / 4; i4 += NWARPS * 32) {
          const float4 z = dither_normal4(seed, (uint32_t)(sbase + 4 * i4) >> 2, (uint32_t)b);
          float4* q = reinterpret_cast<float4*>(rw) + i4;
          float4 v = *q;
          v.x = fmaf(dither, z.x, v.x); v.y = fmaf(dither, z.y, v.y); v.z = fmaf(dither, z.z, v.z); v.w = fmaf(dither, z.w, v.w);
          *q = v;
        }
      } else {
        for (int i = tid + 3; i < span + 4; i += NWARPS * 32) rw[i] = fmaf(dither, dither_normal(seed, sbase + i, b), rw[i]);
      }
      __syncthreads();
    }

    constexpr int PAIRS_PER_WARP = FT / NWARPS / 2;
#pragma unroll 1
    for (int pp = 0; pp < PAIRS_PER_WARP; ++pp) {
      const int fl1 = (warp * PAIRS_PER_WARP + pp) * 2;  // local frame indices fl1, fl1 + 1
      if (f0 + fl1 >= F) break;                          // warp-uniform
      const float* y1 = raw + fl1 * hop;
      const float* y2 = y1 + hop;

      float ar[2][8], ai[2][8];
      // ---- pass 1: butterflies g = lane + 32 h = n0 + 8 n1, inputs n = g + 64 n2 -----------------------
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
          const int q = h + 2 * n2;
          if (q < Q0 || q >= Q1) {
            ar[h][n2] = 0.f;
            ai[h][n2] = 0.f;
          } else {
            const int n = lane + 32 * q;
            const float w = wq[q - Q0], wp = pe * wpq[q - Q0];
            ar[h][n2] = fmaf(-wp, y1[n - 1], w * y1[n]);
            ai[h][n2] = fmaf(-wp, y2[n - 1], w * y2[n]);
          }
        }
        dft8(ar[h], ai[h]);
        const int n1 = (lane >> 3) + 4 * h;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {  // twiddle W64^{n1 k2} = W512^{8 n1 k2}
          float r = ar[h][k2], i = ai[h][k2];
          if (k2 > 0) {
            const float2 t = tw[8 * n1 * k2];
            const float rr = r * t.x - i * t.y;
            i = r * t.y + i * t.x;
            r = rr;
          }
          xch[XS1 * k2 + lane + 32 * h] = make_float2(r, i);
        }
      }
      __syncwarp();
      // ---- pass 2: (n0 = lane % 8, k2 = lane / 8 + 4 h), sum over n1 -----------------------------------
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k2 = (lane >> 3) + 4 * h;
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
          const float2 v = xch[XS1 * k2 + 8 * n1 + n0];
          ar[h][n1] = v.x;
          ai[h][n1] = v.y;
        }
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k2 = (lane >> 3) + 4 * h;
        dft8(ar[h], ai[h]);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {  // twiddle W512^{n0 (k2 + 8 k1)}
          const float2 t = tw[n0 * (k2 + 8 * k1)];
          const float r = ar[h][k1], i = ai[h][k1];
          xch[XS0 * n0 + 8 * k1 + k2] = make_float2(r * t.x - i * t.y, r * t.y + i * t.x);
        }
      }
      __syncwarp();
      // ---- pass 3: j = lane + 32 h = k2 + 8 k1, sum over n0, output Z[j + 64 k0] -----------------------
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
#pragma unroll
        for (int m0 = 0; m0 < 8; ++m0) {
          const float2 v = xch[XS0 * m0 + j];
          ar[h][m0] = v.x;
          ai[h][m0] = v.y;
        }
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        dft8(ar[h], ai[h]);
#pragma unroll
        for (int k0 = 0; k0 < 8; ++k0) xch[j + 64 * k0] = make_float2(ar[h][k0], ai[h][k0]);
      }
      __syncwarp();
      // ---- un-pack the two real spectra and take |.|^2 (sqrt then square, transform.py:205-207) -------
      float2 pw[9];
#pragma unroll
      for (int m = 0; m < 9; ++m) {
        const int k = lane + 32 * m;
        pw[m] = make_float2(0.f, 0.f);
        if (k < NBINS) {
          const float2 z = xch[k], c = xch[(NFFT - k) & (NFFT - 1)];
          const float r1 = 0.5f * (z.x + c.x), i1 = 0.5f * (z.y - c.y);
          const float r2 = 0.5f * (z.y + c.y), i2 = 0.5f * (c.x - z.x);
          const float m1 = sqrtf(r1 * r1 + i1 * i1);
          const float m2 = sqrtf(r2 * r2 + i2 * i2);
          pw[m] = make_float2(m1 * m1, m2 * m2);
        }
      }
      __syncwarp();
#pragma unroll
      for (int m = 0; m < 9; ++m) {
        const int k = lane + 32 * m;
        if (k < NBINS) xch[k] = pw[m];
      }
      __syncwarp();
      // ---- sparse mel projection + log (transform.py:250-254), both frames per shared-memory read ------
      for (int m = lane; m < nfilt; m += 32) {
        const int st = s_mstart[m], cnt = s_mcount[m], off = s_moff[m];
        float a1 = 0.f, a2 = 0.f;
        for (int j = 0; j < cnt; ++j) {
          const float w = s_mw[off + j];
          const float2 pv = xch[st + j];
          a1 = fmaf(w, pv.x, a1);
          a2 = fmaf(w, pv.y, a2);
        }
        tile[m * TILE_LD + fl1] = logf(a1 + 5.9604644775390625e-08f);  // 2^-24
        tile[m * TILE_LD + fl1 + 1] = logf(a2 + 5.9604644775390625e-08f);
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- coalesced write-out: one (filter) row segment of FT frames per warp instruction -------------
    for (int m = warp; m < nfilt; m += NWARPS) {
      const int f = f0 + lane;
      if (lane < FT && f < F) logmel[((size_t)b * nfilt + m) * F + f] = tile[m * TILE_LD + lane];
    }
    // the next iteration's first __syncthreads orders these tile reads before the next item's tile writes
  }
}

__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store_out(__half* p, float v) { *p = __float2half_rn(v); }

constexpr int NORM_WARPS = 8;
constexpr int NORM_REG = 64;  // most values per lane kept in registers => F <= 2048 single pass

// REG = values per lane held in registers (instantiated for 8 / 16 / 24 / 32 / 48 / 64: the launcher picks the smallest
// that covers max(F, out_pitch), so a 751-frame row costs 24 predicated iterations per pass instead of 64)
template <typename OutT, int REG>
__global__ void __launch_bounds__(NORM_WARPS * 32)
normalize_rows_kernel(const float* __restrict__ logmel, const int64_t* __restrict__ lengths, int rows,
                      int nfilt, int F, int hop, float div_guard, OutT* __restrict__ out, int out_pitch,
                      int64_t* __restrict__ seq_len_out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * NORM_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row / nfilt;
  const int64_t len = lengths[b];
  // floor(len / hop) + 1  (PowerSpectrum.get_sequence_length, transform.py:182-184)
  const int64_t q = len >= 0 ? len / hop : -((-len + hop - 1) / hop);
  const int64_t seq = q + 1;
  if (seq_len_out != nullptr && row % nfilt == 0 && lane == 0) seq_len_out[b] = seq;
  const int n = (int)(seq < 0 ? 0 : (seq > F ? F : seq));  // number of valid frames = mask.sum()
  const float* in = logmel + (size_t)row * F;
  OutT* o = out + (size_t)row * out_pitch;

  if (F <= REG * 32 && out_pitch <= REG * 32) {
    float v[REG];
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < REG; ++j) {
      const int t = lane + 32 * j;
      v[j] = (t < n) ? in[t] : 0.f;
      s += (double)v[j];
    }
    s = warp_sum(s);
    const float mean = (float)(s / (double)n);
    double ss = 0.0;
#pragma unroll
    for (int j = 0; j < REG; ++j) {
      const int t = lane + 32 * j;
      const float d = (t < n) ? (v[j] - mean) : 0.f;
      ss += (double)d * (double)d;
    }
    // the reference subtracts the mean from the zero-FILLED tensor (blocks.py:141-145): each of the F - n masked
    // frames contributes mean^2 to the numerator
    ss = warp_sum(ss) + (double)(F - n) * (double)mean * (double)mean;
    const float stdv = (float)sqrt(ss / (double)n);
    const float den = stdv + div_guard;
#pragma unroll
    for (int j = 0; j < REG; ++j) {
      const int t = lane + 32 * j;
      if (t < out_pitch) store_out(o + t, (t < n) ? (v[j] - mean) / den : 0.f);
    }
  } else {
    double s = 0.0;
    for (int t = lane; t < n; t += 32) s += (double)in[t];
    s = warp_sum(s);
    const float mean = (float)(s / (double)n);
    double ss = 0.0;
    for (int t = lane; t < n; t += 32) {
      const float d = in[t] - mean;
      ss += (double)d * (double)d;
    }
    ss = warp_sum(ss) + (double)(F - n) * (double)mean * (double)mean;
    const float den = (float)sqrt(ss / (double)n) + div_guard;
    for (int t = lane; t < out_pitch; t += 32) store_out(o + t, (t < n) ? (in[t] - mean) / den : 0.f);
  }
}

// Same normalisation from the per-(32 frames, filter) partial sums the DFT front-end emits (ts_logmel_dft): no reduction
// pass over the log-mel tensor -- 4 warps per (b, filter) row each re-derive mean / std from <= 64 partials and stream a
// quarter of the row.  sum (x - mean)^2 over the valid frames = S2 - n mean^2; the reference's (0 - mean)^2 term of every
// masked frame is added like in normalize_rows_kernel.
template <typename OutT>
__global__ void __launch_bounds__(128)
normalize_partials_kernel(const float* __restrict__ logmel, const float* __restrict__ partials, int pslots,
                          const int64_t* __restrict__ lengths, int nfilt, int F, int hop, float div_guard,
                          OutT* __restrict__ out, int out_pitch, int64_t* __restrict__ seq_len_out) {
  const int row = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = row / nfilt, m = row - b * nfilt;
  const int64_t len = lengths[b];
  const int64_t q = len >= 0 ? len / hop : -((-len + hop - 1) / hop);
  const int64_t seq = q + 1;
  if (seq_len_out != nullptr && m == 0 && threadIdx.x == 0) seq_len_out[b] = seq;
  const int n = (int)(seq < 0 ? 0 : (seq > F ? F : seq));
  double s1 = 0.0, s2 = 0.0;
  for (int i = lane; i < pslots; i += 32) {
    const float2 v = *reinterpret_cast<const float2*>(partials + (((size_t)b * pslots + i) * nfilt + m) * 2);
    s1 += (double)v.x;
    s2 += (double)v.y;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const float mean = (float)(s1 / (double)n);
  const double md = (double)mean;
  double ss = s2 - 2.0 * md * s1 + (double)n * md * md + (double)(F - n) * md * md;
  if (ss < 0.0) ss = 0.0;
  const float den = (float)sqrt(ss / (double)n) + div_guard;
  const float* in = logmel + (size_t)row * F;
  OutT* o = out + (size_t)row * out_pitch;
  for (int t = threadIdx.x; t < out_pitch; t += 128) store_out(o + t, (t < n) ? (in[t] - mean) / den : 0.f);
  (void)warp;
}

}  // namespace feat
}  // namespace ts

using namespace ts;

extern "C" int ts_feature_normalize_partials(const float* logmel, const float* partials, const int64_t* lengths, int B,
                                             int nfilt, int F, int hop, float div_guard, void* out, int out_dtype,
                                             int out_pitch, int64_t* seq_len_out, void* stream) {
  TS_REQUIRE(logmel && partials && lengths && out, TS_ERR_INVALID, "ts_feature_normalize_partials: null pointer");
  TS_REQUIRE(B > 0 && nfilt > 0 && F > 0 && hop > 0 && out_pitch >= F, TS_ERR_INVALID, "ts_feature_normalize_partials: bad sizes");
  TS_REQUIRE(out_dtype == TS_F32 || out_dtype == TS_BF16 || out_dtype == TS_F16, TS_ERR_INVALID,
             "ts_feature_normalize_partials: bad dtype %d", out_dtype);
  const int rows = B * nfilt, pslots = ceil_div(F, 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == TS_F32)
    feat::normalize_partials_kernel<float><<<rows, 128, 0, st>>>(logmel, partials, pslots, lengths, nfilt, F, hop, div_guard,
                                                                 (float*)out, out_pitch, seq_len_out);
  else if (out_dtype == TS_F16)
    feat::normalize_partials_kernel<__half><<<rows, 128, 0, st>>>(logmel, partials, pslots, lengths, nfilt, F, hop, div_guard,
                                                                  (__half*)out, out_pitch, seq_len_out);
  else
    feat::normalize_partials_kernel<__nv_bfloat16><<<rows, 128, 0, st>>>(logmel, partials, pslots, lengths, nfilt, F, hop,
                                                                         div_guard, (__nv_bfloat16*)out, out_pitch, seq_len_out);
  TS_LAUNCH_CHECK("normalize_partials_kernel");
  return TS_OK;
}

static int logmel_launch(const float* audio, int B, int N, int n_fft, int hop, float preemph,
                         const float* window_full, int win_lo, int win_hi, const float* twiddle,
                         const int32_t* mel_start,
                         const int32_t* mel_count, const int32_t* mel_off, const float* mel_w, int nfilt,
                         int nnz, float* logmel, float dither, unsigned long long seed,
                         const unsigned long long* seed_dev, void* stream) {
  TS_REQUIRE(audio && window_full && twiddle && mel_start && mel_count && mel_off && mel_w && logmel,
             TS_ERR_INVALID, "ts_logmel: null pointer");
  TS_REQUIRE(B > 0 && hop > 0 && nfilt > 0 && nnz > 0, TS_ERR_INVALID, "ts_logmel: bad sizes B=%d hop=%d nfilt=%d", B,
             hop, nfilt);
  TS_REQUIRE(n_fft >= 2 && n_fft <= 8192 && n_fft % 2 == 0, TS_ERR_UNSUPPORTED, "ts_logmel: n_fft=%d is out of range (even, 2..8192)",
             n_fft);
  TS_REQUIRE(N > n_fft / 2, TS_ERR_INVALID,
             "ts_logmel: reflect padding needs N > n_fft/2 (N=%d), same as torch.stft(center=True)", N);
  if (n_fft != feat::NFFT) {   // direct-DFT fallback for every other size
    TS_REQUIRE(0 <= win_lo && win_lo < win_hi && win_hi <= n_fft, TS_ERR_INVALID, "ts_logmel: bad window support [%d,%d)",
               win_lo, win_hi);
    const int F = 1 + N / hop, wlen = win_hi - win_lo, nbins = n_fft / 2 + 1;
    auto smem_for = [&](int gf) { return (size_t)(2 * n_fft + gf * (wlen + nbins)) * sizeof(float); };
    int gf = 8;
    while (gf > 1 && smem_for(gf) > 200 * 1024) gf >>= 1;
    TS_REQUIRE(smem_for(gf) <= 200 * 1024, TS_ERR_UNSUPPORTED, "ts_logmel: n_fft=%d needs %zu bytes of shared memory", n_fft,
               smem_for(gf));
    TS_REQUIRE(B <= 65535, TS_ERR_UNSUPPORTED, "ts_logmel: batch %d too large for the generic n_fft path", B);
    auto kern = gf == 8 ? feat::logmel_generic_kernel<8> : gf == 4 ? feat::logmel_generic_kernel<4>
              : gf == 2 ? feat::logmel_generic_kernel<2> : feat::logmel_generic_kernel<1>;
    TS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(gf)));
    kern<<<dim3(ceil_div(F, gf), B), 256, smem_for(gf), (cudaStream_t)stream>>>(
        audio, N, F, n_fft, hop, preemph, window_full, win_lo, win_hi, reinterpret_cast<const float2*>(twiddle), mel_start,
        mel_count, mel_off, mel_w, nfilt, logmel, dither, seed, seed_dev);
    TS_LAUNCH_CHECK("logmel_generic_kernel");
    return TS_OK;
  }
  TS_REQUIRE(nfilt <= feat::MAX_NFILT && nnz <= feat::MAX_NNZ, TS_ERR_UNSUPPORTED,
             "ts_logmel: filter bank too dense (nfilt=%d nnz=%d, limits %d/%d)", nfilt, nnz, feat::MAX_NFILT,
             feat::MAX_NNZ);
  const int F = 1 + N / hop;
  const feat::SmemLayout L = feat::smem_layout(hop, nfilt, nnz);
  const size_t smem = (size_t)L.total * sizeof(float);
  TS_REQUIRE(smem <= 113 * 1024, TS_ERR_UNSUPPORTED, "ts_logmel: hop=%d needs %zu bytes of shared memory", hop, smem);
  TS_REQUIRE(0 <= win_lo && win_lo < win_hi && win_hi <= n_fft, TS_ERR_INVALID, "ts_logmel: bad window support [%d,%d)",
             win_lo, win_hi);
  // sparse-input FFT when the window support is inside [96, 416) (the reference default: 320 centred in 512)
  const bool centre320 = win_lo >= 96 && win_hi <= 416;
  auto kern = centre320 ? feat::logmel_kernel<true> : feat::logmel_kernel<false>;
  TS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    TS_CUDA(cudaGetDevice(&dev));
    TS_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles_per_row = ceil_div(F, feat::FT);
  const long long items_ll = (long long)B * tiles_per_row;
  TS_REQUIRE(items_ll < (1ll << 31), TS_ERR_UNSUPPORTED, "ts_logmel: too many frame tiles");
  const int num_items = (int)items_ll;
  const int grid = num_items < 2 * num_sms ? num_items : 2 * num_sms;
  kern<<<grid, feat::NWARPS * 32, smem, (cudaStream_t)stream>>>(
      audio, B, N, F, hop, preemph, window_full, reinterpret_cast<const float2*>(twiddle), mel_start, mel_count,
      mel_off, mel_w, nfilt, nnz, logmel, tiles_per_row, num_items, dither, seed, seed_dev);
  TS_LAUNCH_CHECK("logmel_kernel");
  return TS_OK;
}

extern "C" int ts_logmel(const float* audio, int B, int N, int n_fft, int hop, float preemph,
                         const float* window_full, int win_lo, int win_hi, const float* twiddle,
                         const int32_t* mel_start,
                         const int32_t* mel_count, const int32_t* mel_off, const float* mel_w, int nfilt,
                         int nnz, float* logmel, void* stream) {
  return logmel_launch(audio, B, N, n_fft, hop, preemph, window_full, win_lo, win_hi, twiddle, mel_start, mel_count, mel_off,
                       mel_w, nfilt, nnz, logmel, 0.f, 0ull, nullptr, stream);
}

extern "C" int ts_logmel_dither(const float* audio, int B, int N, int n_fft, int hop, float preemph,
                                const float* window_full, int win_lo, int win_hi, const float* twiddle,
                                const int32_t* mel_start, const int32_t* mel_count, const int32_t* mel_off,
                                const float* mel_w, int nfilt, int nnz, float* logmel, float dither,
                                unsigned long long seed, const unsigned long long* seed_dev, void* stream) {
  return logmel_launch(audio, B, N, n_fft, hop, preemph, window_full, win_lo, win_hi, twiddle, mel_start, mel_count, mel_off,
                       mel_w, nfilt, nnz, logmel, dither, seed, seed_dev, stream);
}

extern "C" int ts_feature_normalize(const float* logmel, const int64_t* lengths, int B, int nfilt, int F, int hop,
                                    float div_guard, void* out, int out_dtype, int out_pitch, int64_t* seq_len_out,
                                    void* stream) {
  TS_REQUIRE(logmel && lengths && out, TS_ERR_INVALID, "ts_feature_normalize: null pointer");
  TS_REQUIRE(B > 0 && nfilt > 0 && F > 0 && hop > 0, TS_ERR_INVALID, "ts_feature_normalize: bad sizes");
  TS_REQUIRE(out_pitch >= F, TS_ERR_INVALID, "ts_feature_normalize: out_pitch %d < F %d", out_pitch, F);
  TS_REQUIRE(out_dtype == TS_F32 || out_dtype == TS_BF16 || out_dtype == TS_F16, TS_ERR_INVALID,
             "ts_feature_normalize: bad dtype %d", out_dtype);
  const int rows = B * nfilt;
  const int grid = ceil_div(rows, feat::NORM_WARPS);
  const int need = ceil_div(F > out_pitch ? F : out_pitch, 32);   // values per lane
#define TS_NORM_LAUNCH(T, REG)                                                                                     \
  feat::normalize_rows_kernel<T, REG><<<grid, feat::NORM_WARPS * 32, 0, (cudaStream_t)stream>>>(                  \
      logmel, lengths, rows, nfilt, F, hop, div_guard, (T*)out, out_pitch, seq_len_out)
#define TS_NORM_DISPATCH(T)                       \
  do {                                            \
    if (need <= 8) TS_NORM_LAUNCH(T, 8);          \
    else if (need <= 16) TS_NORM_LAUNCH(T, 16);   \
    else if (need <= 24) TS_NORM_LAUNCH(T, 24);   \
    else if (need <= 32) TS_NORM_LAUNCH(T, 32);   \
    else if (need <= 48) TS_NORM_LAUNCH(T, 48);   \
    else TS_NORM_LAUNCH(T, 64);                   \
  } while (0)
  if (out_dtype == TS_F32) TS_NORM_DISPATCH(float);
  else if (out_dtype == TS_F16) TS_NORM_DISPATCH(__half);
  else TS_NORM_DISPATCH(__nv_bfloat16);
#undef TS_NORM_DISPATCH
#undef TS_NORM_LAUNCH
  TS_LAUNCH_CHECK("normalize_rows_kernel");
  return TS_OK;
}
