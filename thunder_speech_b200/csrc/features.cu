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
constexpr int XBUF = 576;    // floats per exchange array (max index 72*7+63 = 567)
constexpr int XS1 = 72;      // exchange 1: idx = 72*k2 + 8*n1 + n0
constexpr int XS0 = 68;      // exchange 2: idx = 68*n0 + 8*k1 + k2
constexpr int TILE_LD = FT + 1;
constexpr int MAX_NNZ = 2048;
constexpr int MAX_NFILT = 128;

struct SmemLayout {
  int ybuf, tw_re, tw_im, win, mstart, mcount, moff, mw, xch, tile, total;
};

__host__ __device__ inline SmemLayout smem_layout(int hop, int nfilt, int nnz) {
  SmemLayout L;
  int o = 0;
  L.ybuf = o;   o += ((FT - 1) * hop + NFFT + 3) & ~3;
  L.tw_re = o;  o += NFFT;
  L.tw_im = o;  o += NFFT;
  L.win = o;    o += NFFT;
  L.mstart = o; o += (nfilt + 3) & ~3;
  L.mcount = o; o += (nfilt + 3) & ~3;
  L.moff = o;   o += (nfilt + 3) & ~3;
  L.mw = o;     o += (nnz + 3) & ~3;
  L.xch = o;    o += NWARPS * 2 * XBUF;
  L.tile = o;   o += nfilt * TILE_LD;
  L.total = o;
  return L;
}

// kCentre320: window support is [96, 416) (win_length 320 centred in 512) => only FFT input slots
// q in [3, 12] (n = lane + 32 q) are non-zero for every lane, the rest are compile-time zeros.
template <bool kCentre320>
__global__ void __launch_bounds__(NWARPS * 32, 2)
logmel_kernel(const float* __restrict__ audio, int N, int F, int hop, float preemph,
              const float* __restrict__ window_full, const float2* __restrict__ twiddle,
              const int32_t* __restrict__ mel_start, const int32_t* __restrict__ mel_count,
              const int32_t* __restrict__ mel_off, const float* __restrict__ mel_w, int nfilt, int nnz,
              float* __restrict__ logmel) {
  extern __shared__ __align__(16) float smem[];
  const SmemLayout L = smem_layout(hop, nfilt, nnz);
  float* ybuf = smem + L.ybuf;
  float* tw_re = smem + L.tw_re;
  float* tw_im = smem + L.tw_im;
  float* win = smem + L.win;
  int* s_mstart = reinterpret_cast<int*>(smem + L.mstart);
  int* s_mcount = reinterpret_cast<int*>(smem + L.mcount);
  int* s_moff = reinterpret_cast<int*>(smem + L.moff);
  float* s_mw = smem + L.mw;
  float* tile = smem + L.tile;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * FT;
  const float* x = audio + (size_t)b * N;

  // ---- stage tables and the pre-emphasised, reflect-padded audio span of this frame tile -------------
  for (int i = tid; i < NFFT; i += NWARPS * 32) {
    float2 t = twiddle[i];
    tw_re[i] = t.x;
    tw_im[i] = t.y;
    win[i] = window_full[i];
  }
  for (int i = tid; i < nfilt; i += NWARPS * 32) {
    s_mstart[i] = mel_start[i];
    s_mcount[i] = mel_count[i];
    s_moff[i] = mel_off[i];
  }
  for (int i = tid; i < nnz; i += NWARPS * 32) s_mw[i] = mel_w[i];

  const int span = (FT - 1) * hop + NFFT;
  const int s0 = f0 * hop - NFFT / 2;
  for (int i = tid; i < span; i += NWARPS * 32) {
    int s = s0 + i;
    float v = 0.f;
    if (s > -NFFT / 2 - 1 && s < N + NFFT / 2) {
      int r = s < 0 ? -s : (s >= N ? 2 * (N - 1) - s : s);  // reflect (torch.stft center=True)
      // y[0] = x[0]; y[n] = x[n] - preemph * x[n-1]   (PreEmphasisFilter over the whole padded row)
      v = (r >= 1) ? (x[r] - preemph * x[r - 1]) : x[0];
    }
    ybuf[i] = v;
  }
  __syncthreads();

  float* xre = smem + L.xch + warp * 2 * XBUF;
  float* xim = xre + XBUF;

  constexpr int PAIRS_PER_WARP = FT / NWARPS / 2;
#pragma unroll 1
  for (int p = 0; p < PAIRS_PER_WARP; ++p) {
    const int fl1 = (warp * PAIRS_PER_WARP + p) * 2;  // local frame indices fl1, fl1 + 1
    if (f0 + fl1 >= F) break;                        // warp-uniform
    const float* y1 = ybuf + fl1 * hop;
    const float* y2 = y1 + hop;

    float ar[2][8], ai[2][8];
    // ---- pass 1: butterflies g = lane + 32 h = n0 + 8 n1, inputs n = g + 64 n2 ----------------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) {
        const int q = h + 2 * n2;
        if (kCentre320 && (q < 3 || q > 12)) {
          ar[h][n2] = 0.f;
          ai[h][n2] = 0.f;
        } else {
          const int n = lane + 32 * q;
          const float w = win[n];
          ar[h][n2] = w * y1[n];
          ai[h][n2] = w * y2[n];
        }
      }
      dft8(ar[h], ai[h]);
      const int n1 = (lane >> 3) + 4 * h;
#pragma unroll
      for (int k2 = 1; k2 < 8; ++k2) {  // twiddle W64^{n1 k2} = W512^{8 n1 k2}
        const int e = 8 * n1 * k2;
        const float c = tw_re[e], s = tw_im[e];
        const float r = ar[h][k2], i = ai[h][k2];
        ar[h][k2] = r * c - i * s;
        ai[h][k2] = r * s + i * c;
      }
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) {
        xre[XS1 * k2 + lane + 32 * h] = ar[h][k2];
        xim[XS1 * k2 + lane + 32 * h] = ai[h][k2];
      }
    }
    __syncwarp();
    // ---- pass 2: (n0 = lane % 8, k2 = lane / 8 + 4 h), sum over n1 ----------------------------------
    const int n0 = lane & 7;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k2 = (lane >> 3) + 4 * h;
#pragma unroll
      for (int n1 = 0; n1 < 8; ++n1) {
        ar[h][n1] = xre[XS1 * k2 + 8 * n1 + n0];
        ai[h][n1] = xim[XS1 * k2 + 8 * n1 + n0];
      }
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k2 = (lane >> 3) + 4 * h;
      dft8(ar[h], ai[h]);
#pragma unroll
      for (int k1 = 0; k1 < 8; ++k1) {  // twiddle W512^{n0 (k2 + 8 k1)}
        const int e = n0 * (k2 + 8 * k1);
        const float c = tw_re[e], s = tw_im[e];
        const float r = ar[h][k1], i = ai[h][k1];
        xre[XS0 * n0 + 8 * k1 + k2] = r * c - i * s;
        xim[XS0 * n0 + 8 * k1 + k2] = r * s + i * c;
      }
    }
    __syncwarp();
    // ---- pass 3: j = lane + 32 h = k2 + 8 k1, sum over n0, output Z[j + 64 k0] ----------------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
#pragma unroll
      for (int m0 = 0; m0 < 8; ++m0) {
        ar[h][m0] = xre[XS0 * m0 + j];
        ai[h][m0] = xim[XS0 * m0 + j];
      }
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
      dft8(ar[h], ai[h]);
#pragma unroll
      for (int k0 = 0; k0 < 8; ++k0) {
        xre[j + 64 * k0] = ar[h][k0];
        xim[j + 64 * k0] = ai[h][k0];
      }
    }
    __syncwarp();
    // ---- un-pack the two real spectra and take |.|^2 (sqrt then square, transform.py:205-207) ------
    float p1[9], p2[9];
#pragma unroll
    for (int m = 0; m < 9; ++m) {
      const int k = lane + 32 * m;
      p1[m] = 0.f;
      p2[m] = 0.f;
      if (k < NBINS) {
        const int kk = (NFFT - k) & (NFFT - 1);
        const float zr = xre[k], zi = xim[k], cr = xre[kk], ci = xim[kk];
        const float r1 = 0.5f * (zr + cr), i1 = 0.5f * (zi - ci);
        const float r2 = 0.5f * (zi + ci), i2 = 0.5f * (cr - zr);
        const float m1 = sqrtf(r1 * r1 + i1 * i1);
        const float m2 = sqrtf(r2 * r2 + i2 * i2);
        p1[m] = m1 * m1;
        p2[m] = m2 * m2;
      }
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 9; ++m) {
      const int k = lane + 32 * m;
      if (k < NBINS) {
        xre[k] = p1[m];
        xim[k] = p2[m];
      }
    }
    __syncwarp();
    // ---- sparse mel projection + log (transform.py:250-254) -----------------------------------------
    for (int m = lane; m < nfilt; m += 32) {
      const int st = s_mstart[m], cnt = s_mcount[m], off = s_moff[m];
      float a1 = 0.f, a2 = 0.f;
      for (int j = 0; j < cnt; ++j) {
        const float w = s_mw[off + j];
        a1 = fmaf(w, xre[st + j], a1);
        a2 = fmaf(w, xim[st + j], a2);
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
}

__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int NORM_WARPS = 8;
constexpr int NORM_REG = 64;  // values per lane kept in registers => F <= 2048 single pass

template <typename OutT>
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

  if (F <= NORM_REG * 32) {
    float v[NORM_REG];
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NORM_REG; ++j) {
      const int t = lane + 32 * j;
      v[j] = (t < n) ? in[t] : 0.f;
      s += (double)v[j];
    }
    s = warp_sum(s);
    const float mean = (float)(s / (double)n);
    double ss = 0.0;
#pragma unroll
    for (int j = 0; j < NORM_REG; ++j) {
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
    for (int j = 0; j < NORM_REG; ++j) {
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

}  // namespace feat
}  // namespace ts

using namespace ts;

extern "C" int ts_logmel(const float* audio, int B, int N, int n_fft, int hop, float preemph,
                         const float* window_full, int win_lo, int win_hi, const float* twiddle,
                         const int32_t* mel_start,
                         const int32_t* mel_count, const int32_t* mel_off, const float* mel_w, int nfilt,
                         int nnz, float* logmel, void* stream) {
  TS_REQUIRE(audio && window_full && twiddle && mel_start && mel_count && mel_off && mel_w && logmel,
             TS_ERR_INVALID, "ts_logmel: null pointer");
  TS_REQUIRE(B > 0 && hop > 0 && nfilt > 0 && nnz > 0, TS_ERR_INVALID, "ts_logmel: bad sizes B=%d hop=%d nfilt=%d", B,
             hop, nfilt);
  TS_REQUIRE(n_fft == feat::NFFT, TS_ERR_UNSUPPORTED, "ts_logmel: only n_fft=512 is implemented (got %d)", n_fft);
  TS_REQUIRE(N > n_fft / 2, TS_ERR_INVALID,
             "ts_logmel: reflect padding needs N > n_fft/2 (N=%d), same as torch.stft(center=True)", N);
  TS_REQUIRE(nfilt <= feat::MAX_NFILT && nnz <= feat::MAX_NNZ, TS_ERR_UNSUPPORTED,
             "ts_logmel: filter bank too dense (nfilt=%d nnz=%d, limits %d/%d)", nfilt, nnz, feat::MAX_NFILT,
             feat::MAX_NNZ);
  TS_REQUIRE(B <= 65535, TS_ERR_UNSUPPORTED, "ts_logmel: B=%d > 65535", B);
  const int F = 1 + N / hop;
  const feat::SmemLayout L = feat::smem_layout(hop, nfilt, nnz);
  const size_t smem = (size_t)L.total * sizeof(float);
  TS_REQUIRE(smem <= 227 * 1024, TS_ERR_UNSUPPORTED, "ts_logmel: hop=%d needs %zu bytes of shared memory", hop, smem);
  TS_REQUIRE(0 <= win_lo && win_lo < win_hi && win_hi <= n_fft, TS_ERR_INVALID, "ts_logmel: bad window support [%d,%d)",
             win_lo, win_hi);
  // sparse-input FFT when the window support is inside [96, 416) (the reference default: 320 centred in 512)
  const bool centre320 = win_lo >= 96 && win_hi <= 416;
  auto kern = centre320 ? feat::logmel_kernel<true> : feat::logmel_kernel<false>;
  TS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(F, feat::FT), B);
  kern<<<grid, feat::NWARPS * 32, smem, (cudaStream_t)stream>>>(
      audio, N, F, hop, preemph, window_full, reinterpret_cast<const float2*>(twiddle), mel_start, mel_count,
      mel_off, mel_w, nfilt, nnz, logmel);
  TS_LAUNCH_CHECK("logmel_kernel");
  return TS_OK;
}

extern "C" int ts_feature_normalize(const float* logmel, const int64_t* lengths, int B, int nfilt, int F, int hop,
                                    float div_guard, void* out, int out_dtype, int out_pitch, int64_t* seq_len_out,
                                    void* stream) {
  TS_REQUIRE(logmel && lengths && out, TS_ERR_INVALID, "ts_feature_normalize: null pointer");
  TS_REQUIRE(B > 0 && nfilt > 0 && F > 0 && hop > 0, TS_ERR_INVALID, "ts_feature_normalize: bad sizes");
  TS_REQUIRE(out_pitch >= F, TS_ERR_INVALID, "ts_feature_normalize: out_pitch %d < F %d", out_pitch, F);
  TS_REQUIRE(out_dtype == TS_F32 || out_dtype == TS_BF16, TS_ERR_INVALID, "ts_feature_normalize: bad dtype %d",
             out_dtype);
  const int rows = B * nfilt;
  const int grid = ceil_div(rows, feat::NORM_WARPS);
  if (out_dtype == TS_F32) {
    feat::normalize_rows_kernel<float><<<grid, feat::NORM_WARPS * 32, 0, (cudaStream_t)stream>>>(
        logmel, lengths, rows, nfilt, F, hop, div_guard, (float*)out, out_pitch, seq_len_out);
  } else {
    feat::normalize_rows_kernel<__nv_bfloat16><<<grid, feat::NORM_WARPS * 32, 0, (cudaStream_t)stream>>>(
        logmel, lengths, rows, nfilt, F, hop, div_guard, (__nv_bfloat16*)out, out_pitch, seq_len_out);
  }
  TS_LAUNCH_CHECK("normalize_rows_kernel");
  return TS_OK;
}
